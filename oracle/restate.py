"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (torch fp32 on host cores) of the reference's stage-two hot path.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import this file, and only as the checker / reported CPU baseline.

Each function states the reference file:line it follows.  The restatement is
functional (explicit state dicts, no nn.Module) so that it shares no code with
either the reference or the product's module mirrors.

Pinning: `tests/test_oracle_golden.py` checks every function here against the
fixtures in `tests/golden/` which were produced by running the UNMODIFIED
reference through `oracle/ref_harness.py` (`oracle/make_golden.py`, committed).
The reference itself holds no golden vectors for this path.
"""
import math

import torch
from torch.nn import functional as F


def leaky(x):
    return F.leaky_relu(x, 0.2)


# ------------------------------------------------------------------ Audio2Mel
def audio2mel(audio, mel_basis, window, n_fft=1024, hop_length=256):
    """featuresynth/feature/feature.py:39-59 (Audio2Mel.forward).

    audio (B,1,N) -> right zero-pad (n_fft-hop)//2 -> STFT(center=False, one-sided)
    -> magnitude -> mel_basis @ mag -> log10(clamp(., 1e-5)).  (B, n_mels, F),
    F = (N + p - n_fft)//hop + 1.
    """
    p = (n_fft - hop_length) // 2
    a = F.pad(audio, (0, p)).squeeze(1)
    frames = a.unfold(-1, n_fft, hop_length)                     # (B, F, n_fft)
    spec = torch.fft.rfft(frames * window, dim=-1)               # (B, F, bins)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2).transpose(1, 2)
    mel = torch.matmul(mel_basis, mag)
    return torch.log10(torch.clamp(mel, min=1e-5))


# ------------------------------------------------------------- residual blocks
def residual_atom(x, w1, b1, w2, b2, dilation):
    """featuresynth/util/modules.py:350-388 (ResidualAtom.forward):
    x + leaky(conv_k3_pad1(leaky(conv_k3_dil_d_pad_d(x))))."""
    y = leaky(F.conv1d(x, w1, b1, dilation=dilation, padding=dilation))
    y = leaky(F.conv1d(y, w2, b2, padding=1))
    return x + y


def residual_stack(x, sd, prefix, dilations=(1, 3, 9)):
    """featuresynth/util/modules.py:391-405 (ResidualStack)."""
    for a, d in enumerate(dilations):
        x = residual_atom(
            x,
            sd[f"{prefix}.main.{a}.main.0.weight"], sd[f"{prefix}.main.{a}.main.0.bias"],
            sd[f"{prefix}.main.{a}.main.1.weight"], sd[f"{prefix}.main.{a}.main.1.bias"],
            d)
    return x


# ------------------------------------------------------------- MelGanGenerator
MELGAN_UPSAMPLERS = ((3, 5, 8, 4), (6, 8, 8, 4), (9, 11, 2, 1), (12, 14, 2, 1))


def melgan_generator(x, sd):
    """featuresynth/generator/full.py:16-50 (MelGanGenerator; weight_norm is the
    identity there, full.py:12-13).  x (B,128,T) -> (B,1,256T)."""
    x = F.pad(x, (3, 3), mode="reflect")
    x = leaky(F.conv1d(x, sd["main.1.weight"], sd["main.1.bias"]))
    for ct, st, s, p in MELGAN_UPSAMPLERS:
        x = leaky(F.conv_transpose1d(
            x, sd[f"main.{ct}.weight"], sd[f"main.{ct}.bias"], stride=s, padding=p))
        x = residual_stack(x, sd, f"main.{st}")
    x = F.conv1d(x, sd["main.15.weight"], sd["main.15.bias"], padding=3)
    return torch.tanh(x)


def melgan_generator_state(seed, in_channels=128):
    """Random-init weights of MelGanGenerator as experiment/init.py:3-9 leaves
    them (weight ~ N(0, 0.02), bias = 0), drawn from a *numpy* RandomState so
    that the fixture generator, the oracle tests and the GPU tests regenerate
    bit-identical tensors on any box."""
    import numpy as np
    rs = np.random.RandomState(seed)

    def w(*shape):
        return torch.from_numpy((rs.standard_normal(shape) * 0.02).astype(np.float32))

    sd = {}
    sd["main.1.weight"] = w(512, in_channels, 7)
    sd["main.1.bias"] = torch.zeros(512)
    chans = {3: (512, 256, 16), 6: (256, 128, 16), 9: (128, 64, 4), 12: (64, 32, 4)}
    for ct, st, s, p in MELGAN_UPSAMPLERS:
        cin, cout, k = chans[ct]
        sd[f"main.{ct}.weight"] = w(cin, cout, k)
        sd[f"main.{ct}.bias"] = torch.zeros(cout)
        for a in range(3):
            for c in range(2):
                sd[f"main.{st}.main.{a}.main.{c}.weight"] = w(cout, cout, 3)
                sd[f"main.{st}.main.{a}.main.{c}.bias"] = torch.zeros(cout)
    sd["main.15.weight"] = w(1, 32, 7)
    sd["main.15.bias"] = torch.zeros(1)
    return sd


def randomize_biases(sd, seed, std=0.01):
    """Non-zero biases so parity tests exercise the bias path (the reference
    init zeroes them; trained checkpoints do not)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    out = dict(sd)
    for k in sorted(sd):
        if k.endswith("bias"):
            out[k] = torch.from_numpy(
                (rs.standard_normal(tuple(sd[k].shape)) * std).astype(np.float32))
    return out


# -------------------------------------------------------------- discriminators
FULL_DISC_LAYERS = (
    # (index, stride, padding, groups)   featuresynth/discriminator/full.py:13-20
    (0, 1, 7, 1), (1, 4, 20, 4), (2, 4, 20, 16), (3, 4, 20, 64), (4, 4, 20, 256),
    (5, 1, 2, 1),
)


def full_discriminator(x, sd, prefix="disc"):
    """featuresynth/discriminator/full.py:34-40 (FullDiscriminator.forward)."""
    feats = []
    for i, s, p, g in FULL_DISC_LAYERS:
        x = leaky(F.conv1d(x, sd[f"{prefix}.main.{i}.weight"],
                           sd[f"{prefix}.main.{i}.bias"], stride=s, padding=p, groups=g))
        feats.append(x)
    j = F.conv1d(x, sd[f"{prefix}.judge.weight"], sd[f"{prefix}.judge.bias"], padding=1)
    return feats, j


def melgan_discriminator(x, sd, scales=2):
    """featuresynth/discriminator/melgan.py:13-27: ONE shared FullDiscriminator
    at 3 scales, avg_pool1d(k4, s2, p2) (count_include_pad=True) between."""
    features, judgements = [], []
    f, j = full_discriminator(x, sd)
    features.append(f)
    judgements.append(j)
    for _ in range(scales):
        x = F.avg_pool1d(x, kernel_size=4, stride=2, padding=2)
        f, j = full_discriminator(x, sd)
        features.append(f)
        judgements.append(j)
    return features, judgements


def melgan_discriminator_state(seed):
    import numpy as np
    rs = np.random.RandomState(seed)

    def w(*shape):
        return torch.from_numpy((rs.standard_normal(shape) * 0.02).astype(np.float32))

    shapes = {0: (16, 1, 15), 1: (64, 4, 41), 2: (256, 4, 41), 3: (1024, 4, 41),
              4: (1024, 4, 41), 5: (1024, 1024, 5)}
    sd = {}
    for i, shp in shapes.items():
        sd[f"disc.main.{i}.weight"] = w(*shp)
        sd[f"disc.main.{i}.bias"] = torch.zeros(shp[0])
    sd["disc.judge.weight"] = w(1, 1024, 3)
    sd["disc.judge.bias"] = torch.zeros(1)
    return sd


# ---------------------------------------------------------------------- losses
def hinge_generator_loss(j):
    """featuresynth/loss/loss.py:9-10."""
    return (-j).mean()


def hinge_discriminator_loss(r_j, f_j):
    """featuresynth/loss/loss.py:17-18."""
    return (F.relu(1 - r_j) + F.relu(1 + f_j)).mean()


def least_squares_generator_loss(j):
    """featuresynth/loss/loss.py:5-6."""
    return 0.5 * ((j - 1) ** 2).mean()


def least_squares_disc_loss(r_j, f_j):
    """featuresynth/loss/loss.py:13-14."""
    return 0.5 * (((r_j - 1) ** 2).mean() + (f_j ** 2).mean())


def mel_gan_disc_loss(real_judgements, fake_judgements,
                      gan_loss=hinge_discriminator_loss):
    """featuresynth/loss/loss.py:21-25."""
    return sum(gan_loss(r, f) for r, f in zip(real_judgements, fake_judgements))


def mel_gan_feature_loss(real_features, fake_features):
    """featuresynth/loss/loss.py:28-65: sum over discriminators and layers of
    (1/n_disc)(1/n_layers) * mean |r - f|."""
    loss = 0
    nd = 1 / len(real_features)
    for r_group, f_group in zip(real_features, fake_features):
        nl = 1 / len(r_group)
        for r_f, f_f in zip(r_group, f_group):
            loss = loss + (nl * nd) * F.l1_loss(r_f, f_f)
    return loss


def mel_gan_gen_loss(real_features, fake_features, real_judgements,
                     fake_judgements, gan_loss=hinge_generator_loss,
                     feature_loss_weight=10):
    """featuresynth/loss/loss.py:68-79."""
    j_loss = sum(gan_loss(f) for _, f in zip(real_judgements, fake_judgements))
    return j_loss + feature_loss_weight * mel_gan_feature_loss(
        real_features, fake_features)


# ------------------------------------------------- FFT multiscale band split/merge
def fft_frequency_decompose(x, min_size):
    """featuresynth/audio/transform.py:50-82: ortho rFFT; band of size S keeps
    bins [S/4, S/2] (the lowest keeps [0, S/2]); ortho irFFT at length S."""
    coeffs = torch.fft.rfft(x, norm="ortho")
    out = {}
    size = min_size
    while size <= x.shape[-1]:
        sl = coeffs[..., :size // 2 + 1].clone()
        if size > min_size:
            sl[..., :size // 4] = 0
        out[size] = torch.fft.irfft(sl, n=size, norm="ortho")
        size *= 2
    return out


def fft_resample(x, desired_size, is_lowest_band):
    """featuresynth/audio/transform.py:85-104."""
    coeffs = torch.fft.rfft(x, norm="ortho")
    n = coeffs.shape[-1]
    new = torch.zeros(x.shape[0], x.shape[1], desired_size // 2 + 1,
                      dtype=coeffs.dtype)
    if is_lowest_band:
        new[..., :n] = coeffs
    else:
        new[..., n // 2:n] = coeffs[..., n // 2:]
    return torch.fft.irfft(new, n=desired_size, norm="ortho")


def fft_frequency_recompose(d, desired_size):
    """featuresynth/audio/transform.py:107-115."""
    first = min(d.keys())
    return sum(fft_resample(b, desired_size, s == first) for s, b in d.items())


# -------------------------------------------------------- Morlet filter bank ops
def filterbank_convolve(x, bank):
    """zounds FilterBank.convolve as used at discriminator/multiscale.py:112:
    conv1d(x(B,1,L), bank(n,1,k), padding=k//2) -> (B,n,L+1) for even k."""
    k = bank.shape[-1]
    return F.conv1d(x.view(-1, 1, x.shape[-1]), bank, padding=k // 2)


def filterbank_transposed_convolve(x, bank):
    """zounds FilterBank.transposed_convolve as used at generator/multiscale.py:91:
    conv_transpose1d(x(B,n,L+1), bank(n,1,k), padding=k//2) -> (B,1,L)."""
    k = bank.shape[-1]
    return F.conv_transpose1d(x, bank, padding=k // 2)


# ------------------------------------------------ FilterBankMultiScaleGenerator
# per band (largest first): LearnedUpSample strides, generator/multiscale.py:113-140
FB_STRIDES = ((4, 4, 4, 4), (4, 4, 4, 2), (4, 4, 2, 2), (4, 2, 2, 2), (2, 2, 2, 2))


def fb_band_sizes(output_size):
    """generator/multiscale.py:110"""
    import numpy as np
    return [int(2 ** (np.log2(output_size) - i)) for i in range(5)]


def fb_banks(samplerate=22050, kernel_size=128, n_bands=128):
    """The five fixed Morlet banks of generator/multiscale.py:113-164 (rates sr, sr/2 ... sr/16,
    each spanning [nyquist/2, nyquist] of its own rate, the lowest [0, nyquist])."""
    from oracle import bases
    out = []
    for i in range(5):
        sr = bases.SampleRate(samplerate) * (2 ** i)
        start = 0 if i == 4 else sr.nyquist / 2
        scale = bases.LinearScale(bases.FrequencyBand(start, sr.nyquist), n_bands)
        out.append(torch.from_numpy(bases.morlet_filter_bank(sr, kernel_size, scale, 0.05))
                   .view(n_bands, 1, kernel_size))
    return out


def fb_generator_state(seed, output_size):
    import numpy as np
    rs = np.random.RandomState(seed)
    sd = {}
    for size, strides in zip(fb_band_sizes(output_size), FB_STRIDES):
        sd[f"channel_{size}.main.0.0.weight"] = torch.from_numpy(
            (rs.standard_normal((128, 128, 7)) * 0.02).astype(np.float32))
        sd[f"channel_{size}.main.0.0.bias"] = torch.from_numpy(
            (rs.standard_normal((128,)) * 0.01).astype(np.float32))
        for li, s_ in enumerate(strides):
            sd[f"channel_{size}.main.{li + 1}.conv.weight"] = torch.from_numpy(
                (rs.standard_normal((128, 128, 2 * s_)) * 0.02).astype(np.float32))
    return sd


def filterbank_multiscale_generator(x, sd, banks, output_size, recompose=False):
    """generator/multiscale.py:86-92 (channel generator) and 166-178 (multi-scale forward);
    LearnedUpSample = ConvTranspose1d(k=2s, stride s, padding s//2, bias=False) + leaky,
    util/modules.py:168-188."""
    results = {}
    for size, strides, bank in zip(fb_band_sizes(output_size), FB_STRIDES, banks):
        h = leaky(F.conv1d(x, sd[f"channel_{size}.main.0.0.weight"],
                           sd[f"channel_{size}.main.0.0.bias"], padding=3))
        for li, s_ in enumerate(strides):
            h = leaky(F.conv_transpose1d(h, sd[f"channel_{size}.main.{li + 1}.conv.weight"],
                                         stride=s_, padding=s_ // 2))
        h = F.pad(h, (0, 1))
        results[size] = filterbank_transposed_convolve(h, bank)
    if recompose:
        up = output_size // x.shape[-1]
        return fft_frequency_recompose(results, x.shape[-1] * up)
    return results


# -------------------------------------------- FilterBankMultiScaleDiscriminator
# per band (largest first): strides of the four k7 convs, discriminator/multiscale.py:141-175
FB_DISC_STRIDES = ((4, 4, 4, 4), (4, 4, 4, 2), (4, 4, 2, 2), (4, 2, 2, 2), (2, 2, 2, 2))


def fb_discriminator_state(seed, input_size, conditioning_channels=128):
    import numpy as np
    rs = np.random.RandomState(seed)

    def w(*shape):
        return torch.from_numpy((rs.standard_normal(shape) * 0.02).astype(np.float32))

    def b(n):
        return torch.from_numpy((rs.standard_normal((n,)) * 0.01).astype(np.float32))

    sd = {}
    for size in fb_band_sizes(input_size):
        for i in range(4):
            sd[f"channel_{size}.main.{i}.weight"] = w(128, 128, 7)
            sd[f"channel_{size}.main.{i}.bias"] = b(128)
        for i in range(3):
            cin = 128 + conditioning_channels if i == 0 else 128
            sd[f"channel_{size}.mj.{i}.weight"] = w(128, cin, 3)
            sd[f"channel_{size}.mj.{i}.bias"] = b(128)
        sd[f"channel_{size}.judge.weight"] = w(1, 128, 3)
        sd[f"channel_{size}.judge.bias"] = b(1)
    for i in range(3):
        cin = 5 * 128 + conditioning_channels if i == 0 else 512
        sd[f"final.{i}.weight"] = w(512, cin, 3)
        sd[f"final.{i}.bias"] = b(512)
    sd["judge.weight"] = w(1, 512, 3)
    sd["judge.bias"] = b(1)
    return sd


def filterbank_multiscale_discriminator(bands, feat, sd, banks, input_size,
                                        conditioning_channels=128):
    """discriminator/multiscale.py:109-126 (per-band) and 212-252 (head), decompose=False:
    `bands` is {size: (B,1,size)}.  Returns (features: 5 x [7] + [3], judgements: 6)."""
    features, channels, judgements = [], [], []
    for size, strides, bank in zip(fb_band_sizes(input_size), FB_DISC_STRIDES, banks):
        x = bands[size]
        f = []
        x = filterbank_convolve(x, bank)[:, :, :x.shape[-1]]
        for i, s_ in enumerate(strides):
            x = leaky(F.conv1d(x, sd[f"channel_{size}.main.{i}.weight"],
                               sd[f"channel_{size}.main.{i}.bias"], stride=s_, padding=3))
            f.append(x)
        if conditioning_channels > 0:
            x = torch.cat([x, feat], dim=1)
        for i in range(3):
            x = leaky(F.conv1d(x, sd[f"channel_{size}.mj.{i}.weight"],
                               sd[f"channel_{size}.mj.{i}.bias"], padding=1))
            f.append(x)
        j = F.conv1d(x, sd[f"channel_{size}.judge.weight"], sd[f"channel_{size}.judge.bias"],
                     padding=1)
        features.append(f)
        channels.append(x)
        judgements.append(j)
    x = torch.cat(channels, dim=1)
    if conditioning_channels > 0:
        up = F.interpolate(feat, size=x.shape[-1])      # F.upsample default: nearest
        x = torch.cat([x, up], dim=1)
    final = []
    for i in range(3):
        x = leaky(F.conv1d(x, sd[f"final.{i}.weight"], sd[f"final.{i}.bias"], padding=1))
        final.append(x)
    features.append(final)
    judgements.append(F.conv1d(x, sd["judge.weight"], sd["judge.bias"], padding=1))
    return features, judgements


# ------------------------------------------- official MelGAN pair (realmelgan.py)
def _wn(sd, name):
    """legacy torch weight_norm (dim 0): w = g * v / ||v||, norm over all dims but 0."""
    v, g = sd[name + ".weight_v"], sd[name + ".weight_g"]
    return g * v / v.reshape(v.shape[0], -1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))


REAL_RATIOS = (8, 8, 2, 2)


def realmelgan_generator_layout(n_residual_layers=3):
    """model indices of experiment/realmelgan.py:48-89: [(kind, index, arg)]"""
    layout, idx = [("first", 1, None)], 2
    for r in REAL_RATIOS:
        layout.append(("up", idx + 1, r))
        idx += 2
        for j in range(n_residual_layers):
            layout.append(("res", idx, 3 ** j))
            idx += 1
    layout.append(("last", idx + 2, None))
    return layout


def realmelgan_generator_state(seed, input_size=128, ngf=32, n_residual_layers=3):
    """weight_g / weight_v / bias tensors in the reference's state-dict naming; v ~ N(0, 0.02),
    g perturbed around ||v|| (weight_norm initialises g = ||v||), small biases."""
    import numpy as np
    rs = np.random.RandomState(seed)
    sd = {}

    def wn(name, shape):
        v = torch.from_numpy((rs.standard_normal(shape) * 0.02).astype(np.float32))
        nrm = v.reshape(shape[0], -1).norm(dim=1)
        g = nrm * torch.from_numpy((1.0 + 0.1 * rs.standard_normal((shape[0],))).astype(np.float32))
        sd[name + ".bias"] = None      # placeholder to keep reference key order: bias, g, v
        sd[name + ".weight_g"] = g.view(-1, *([1] * (len(shape) - 1)))
        sd[name + ".weight_v"] = v

    mult = 16
    for kind, idx, arg in realmelgan_generator_layout(n_residual_layers):
        if kind == "first":
            wn(f"model.{idx}", (mult * ngf, input_size, 7))
            sd[f"model.{idx}.bias"] = torch.from_numpy((rs.standard_normal((mult * ngf,)) * 0.01).astype(np.float32))
        elif kind == "up":
            cin, cout = mult * ngf, mult * ngf // 2
            wn(f"model.{idx}", (cin, cout, 2 * arg))
            sd[f"model.{idx}.bias"] = torch.from_numpy((rs.standard_normal((cout,)) * 0.01).astype(np.float32))
            mult //= 2
        elif kind == "res":
            dim = mult * ngf
            for sub, k in (("block.2", 3), ("block.4", 1), ("shortcut", 1)):
                wn(f"model.{idx}.{sub}", (dim, dim, k))
                sd[f"model.{idx}.{sub}.bias"] = torch.from_numpy((rs.standard_normal((dim,)) * 0.01).astype(np.float32))
        else:
            wn(f"model.{idx}", (1, ngf, 7))
            sd[f"model.{idx}.bias"] = torch.from_numpy((rs.standard_normal((1,)) * 0.01).astype(np.float32))
    return sd


def realmelgan_generator(x, sd, n_residual_layers=3):
    """experiment/realmelgan.py:32-89: official MelGAN generator (weight norm, reflection pads,
    pre-activation ResnetBlocks with 1x1 conv shortcuts)."""
    for kind, idx, arg in realmelgan_generator_layout(n_residual_layers):
        n = f"model.{idx}"
        if kind == "first":
            x = F.conv1d(F.pad(x, (3, 3), mode="reflect"), _wn(sd, n), sd[n + ".bias"])
        elif kind == "up":
            r = arg
            x = F.conv_transpose1d(leaky(x), _wn(sd, n), sd[n + ".bias"], stride=r,
                                   padding=r // 2 + r % 2, output_padding=r % 2)
        elif kind == "res":
            d = arg
            h = F.conv1d(F.pad(leaky(x), (d, d), mode="reflect"), _wn(sd, n + ".block.2"),
                         sd[n + ".block.2.bias"], dilation=d)
            h = F.conv1d(leaky(h), _wn(sd, n + ".block.4"), sd[n + ".block.4.bias"])
            x = F.conv1d(x, _wn(sd, n + ".shortcut"), sd[n + ".shortcut.bias"]) + h
        else:
            x = torch.tanh(F.conv1d(F.pad(leaky(x), (3, 3), mode="reflect"), _wn(sd, n), sd[n + ".bias"]))
    return x


# (name, cin, cout, k, stride, pad, groups) of NLayerDiscriminator(ndf=16, n_layers=4, factor=4)
REAL_DISC_LAYERS = (("layer_0.1", 1, 16, 15, 1, 0, 1), ("layer_1.0", 16, 64, 41, 4, 20, 4),
                    ("layer_2.0", 64, 256, 41, 4, 20, 16), ("layer_3.0", 256, 1024, 41, 4, 20, 64),
                    ("layer_4.0", 1024, 1024, 41, 4, 20, 256), ("layer_5.0", 1024, 1024, 5, 1, 2, 1),
                    ("layer_6", 1024, 1, 3, 1, 1, 1))


def realmelgan_discriminator_state(seed, num_D=3):
    import numpy as np
    rs = np.random.RandomState(seed)
    sd = {}
    for i in range(num_D):
        for name, cin, cout, k, s_, p_, g_ in REAL_DISC_LAYERS:
            shape = (cout, cin // g_, k)
            v = torch.from_numpy((rs.standard_normal(shape) * 0.02).astype(np.float32))
            nrm = v.reshape(cout, -1).norm(dim=1)
            pre = f"model.disc_{i}.model.{name}"
            sd[pre + ".bias"] = torch.from_numpy((rs.standard_normal((cout,)) * 0.01).astype(np.float32))
            sd[pre + ".weight_g"] = (nrm * torch.from_numpy(
                (1.0 + 0.1 * rs.standard_normal((cout,))).astype(np.float32))).view(-1, 1, 1)
            sd[pre + ".weight_v"] = v
    return sd


def realmelgan_discriminator(x, sd, num_D=3):
    """experiment/realmelgan.py:92-181: three INDEPENDENT NLayerDiscriminators on x, pooled x
    (AvgPool1d(4, 2, 1, count_include_pad=False)); features = all but the last map."""
    features, judgements = [], []
    for i in range(num_D):
        h, res = x, []
        for name, cin, cout, k, s_, p_, g_ in REAL_DISC_LAYERS:
            pre = f"model.disc_{i}.model.{name}"
            if name == "layer_0.1":
                h = F.pad(h, (7, 7), mode="reflect")
            h = F.conv1d(h, _wn(sd, pre), sd[pre + ".bias"], stride=s_, padding=p_, groups=g_)
            if name != "layer_6":
                h = leaky(h)
            res.append(h)
        features.append(res[:-1])
        judgements.append(res[-1])
        x = F.avg_pool1d(x, 4, stride=2, padding=1, count_include_pad=False)
    return features, judgements


def real_mel_gan_feature_loss(real_features, fake_features):
    """experiment/realmelgan.py:185-202: fixed weight (1/3)(4/5) per feature map."""
    wt = (1 / 3) * (4.0 / 5)
    loss = 0
    for r_group, f_group in zip(real_features, fake_features):
        for r_f, f_f in zip(r_group, f_group):
            loss = loss + wt * F.l1_loss(r_f, f_f)
    return loss


# ---------------------------------------------------------------- FLOP counting
MELGAN_FLOP_PER_SAMPLE = 409536  # SURVEY.md App. A.1 (2 x MAC, conv/convT only)


def melgan_generator_flops(batch, frames):
    return MELGAN_FLOP_PER_SAMPLE * batch * frames * 256


# ---------------------------------------------------------------------------------------
# FilterBankExperiment pair: FilterBankGenerator (generator/filterbank.py:93-128) and
# FilterBankDiscriminator (discriminator/filterbank.py:114-202) with
# LowResSpectrogramDiscriminator (util/modules.py:275-344) over ONE fixed 511-tap, 128-band
# linear-scale Morlet bank (experiment/filterbank.py:34-45)
# ---------------------------------------------------------------------------------------
def filterbank_experiment_bank(samplerate=22050, kernel_size=511, n_bands=128):
    """experiment/filterbank.py:34-45: LinearScale(FrequencyBand(20, nyquist - 20), 128), 511 taps,
    scaling factor 0.9, unit-norm filters -> (n_bands, 1, kernel_size)"""
    from oracle import bases
    sr = bases.SampleRate(samplerate)
    scale = bases.LinearScale(bases.FrequencyBand(20, sr.nyquist - 20), n_bands)
    return torch.from_numpy(bases.morlet_filter_bank(sr, kernel_size, scale, 0.9)).view(
        n_bands, 1, kernel_size)


def filterbank_generator_state(seed, in_size=32, out_size=8192, in_channels=128, n_bands=128):
    import math
    import numpy as np
    rs = np.random.RandomState(seed)
    n_layers = int(math.log(out_size, 2) - math.log(in_size, 2))
    sd = {}
    for i in range(n_layers):
        cin = in_channels if i == 0 else 256
        sd[f"main.main.{i}.conv.weight"] = torch.from_numpy(
            (rs.standard_normal((cin, 256, 8)) * 0.02).astype(np.float32))
    sd["to_frames.weight"] = torch.from_numpy(
        (rs.standard_normal((n_bands, 256, 7)) * 0.02).astype(np.float32))
    sd["to_frames.bias"] = torch.from_numpy((rs.standard_normal((n_bands,)) * 0.01).astype(np.float32))
    return sd


def filterbank_generator(x, sd, bank):
    """generator/filterbank.py:122-128: UpsamplingStack of LearnedUpSample(C, 256, 8, 2) =
    ConvTranspose1d(k 8, stride 2, padding 3, bias=False) + leaky (util/modules.py:168-188),
    to_frames Conv1d(256, n_bands, 7, 1, 3), filter_bank.transposed_convolve"""
    h = x
    i = 0
    while f"main.main.{i}.conv.weight" in sd:
        h = leaky(F.conv_transpose1d(h, sd[f"main.main.{i}.conv.weight"], stride=2, padding=3))
        i += 1
    h = F.conv1d(h, sd["to_frames.weight"], sd["to_frames.bias"], padding=3)
    return F.conv_transpose1d(h, bank, padding=bank.shape[-1] // 2)


RSFB_CONVT = ((0, 128, 512, 7, 1, 3), (2, 512, 256, 16, 8, 4), (5, 256, 256, 16, 8, 4),
              (8, 256, 256, 4, 2, 1), (11, 256, 256, 4, 2, 1))
RSFB_STACKS = (4, 7, 10, 13)


def resstack_filterbank_generator_state(seed, in_channels=128):
    """state dict of ResidualStackFilterBankGenerator(add_weight_norm=True)
    (generator/filterbank.py:8-66): legacy weight_norm layout bias / weight_g / weight_v; for a
    ConvTranspose1d the norm runs over dim 0 (in_channels): weight_g (cin, 1, 1)"""
    import numpy as np
    rs = np.random.RandomState(seed)

    def t(*shape, std=0.02):
        return torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))

    sd = {}

    def wn(prefix, v):
        sd[prefix + ".bias"] = None                 # placeholder keeps the key order
        sd[prefix + ".weight_g"] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1) * \
            torch.from_numpy((1.0 + 0.1 * rs.standard_normal((v.shape[0], 1, 1))).astype(np.float32))
        sd[prefix + ".weight_v"] = v

    order = []
    for idx, cin, cout, k, s, p in RSFB_CONVT:
        order.append(("convt", idx, cin, cout, k))
    layers = {idx: (cin, cout, k) for idx, cin, cout, k, s, p in RSFB_CONVT}
    for idx in range(14):
        if idx in layers:
            cin, cout, k = layers[idx]
            wn(f"main.{idx}", t(cin, cout, k))
            sd[f"main.{idx}.bias"] = t(cout, std=0.01)
        elif idx in RSFB_STACKS:
            for a in range(3):
                for c in range(2):
                    wn(f"main.{idx}.main.{a}.main.{c}", t(256, 256, 3))
                    sd[f"main.{idx}.main.{a}.main.{c}.bias"] = t(256, std=0.01)
    for name in ("to_frames", "to_noise"):
        wn(name, t(256, 128, 7))
        sd[name + ".bias"] = t(128, std=0.01)
    return sd


def resstack_filterbank_generator(x, sd, bank, raw_noise):
    """generator/filterbank.py:68-90; raw_noise (1,1,256*T) stands for the torch.normal draw"""
    def w(prefix):
        return _wn(sd, prefix)

    h = x
    for idx, cin, cout, k, s, p in RSFB_CONVT:
        h = leaky(F.conv_transpose1d(h, w(f"main.{idx}"), sd[f"main.{idx}.bias"], stride=s, padding=p))
        stack = {0: None, 2: 4, 5: 7, 8: 10, 11: 13}[idx]
        if stack is not None:
            for a, d in enumerate((1, 3, 9)):
                pre = f"main.{stack}.main.{a}.main"
                h = residual_atom(h, w(pre + ".0"), sd[pre + ".0.bias"], w(pre + ".1"),
                                  sd[pre + ".1.bias"], d)
    noise = F.conv_transpose1d(h, w("to_noise"), sd["to_noise.bias"], padding=3)
    filtered = F.conv1d(raw_noise.view(-1, 1, raw_noise.shape[-1]), bank, padding=bank.shape[-1] // 2)
    noise = (noise * filtered).sum(dim=1, keepdim=True)
    harmonic = F.conv_transpose1d(h, w("to_frames"), sd["to_frames.bias"], padding=3)
    harmonic = F.conv_transpose1d(harmonic, bank, padding=bank.shape[-1] // 2)
    return harmonic + noise


def _lowres_channels(freq_bins, max_channels, n_layers, conditioning_channels):
    import math
    out = []
    lc = math.log2(freq_bins)
    for i in range(n_layers):
        cin = int(min(max_channels, 2 ** (i + lc))) + (conditioning_channels if i == 0 else 0)
        cout = int(min(max_channels, 2 ** (i + lc + 1)))
        out.append((cin, cout))
    return out


FBD_MAIN = ((None, 256), (256, 256), (256, 512), (512, 512), (512, 1024), (1024, 1024),
            (1024, 1024), (1024, 1024))
FBD_LOWRES = {"medium_res": (128, 128, 16, 1024), "low_res": (32, 32, 4, 512)}


def filterbank_discriminator_state(seed, n_bands=128, conditioning_channels=0):
    import math
    import numpy as np
    rs = np.random.RandomState(seed)

    def t(*shape, std=0.02):
        return torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))

    sd = {}
    for i, (cin, cout) in enumerate(FBD_MAIN):
        cin = n_bands + conditioning_channels if cin is None else cin
        sd[f"main.{i}.weight"] = t(cout, cin, 7)
        sd[f"main.{i}.bias"] = t(cout, std=0.01)
    sd["judge.weight"] = t(1, 1024, 3)
    sd["judge.bias"] = t(1, std=0.01)
    for name, (fb, ts, nj, mx) in FBD_LOWRES.items():
        n_layers = int(math.log(ts, 2) - math.log(nj, 2))
        chans = _lowres_channels(fb, mx, n_layers, conditioning_channels)
        for i, (cin, cout) in enumerate(chans):
            sd[f"{name}.stack.main.{i}.weight"] = t(cout, cin, 7)
            sd[f"{name}.stack.main.{i}.bias"] = t(cout, std=0.01)
        sd[f"{name}.judge.weight"] = t(1, chans[-1][1], 3)
        sd[f"{name}.judge.bias"] = t(1, std=0.01)
    return sd


def _lowres_discriminator(a, feat, sd, name, conditioning_channels):
    """util/modules.py:315-344"""
    fb, ts, nj, mx = FBD_LOWRES[name]
    batch, channels, time = a.shape
    cw, tw = channels // fb, time // ts
    low = F.avg_pool2d(F.relu(a)[:, None, :, :], (cw, tw)).view(-1, fb, ts)
    if conditioning_channels > 0:
        if feat.shape[-1] < low.shape[-1]:
            feat = F.interpolate(feat, size=low.shape[-1])
        elif feat.shape[-1] > low.shape[-1]:
            feat = F.avg_pool1d(feat, feat.shape[-1] // low.shape[-1])
        low = torch.cat([low, feat], dim=1)
    features = []
    i = 0
    while f"{name}.stack.main.{i}.weight" in sd:
        low = leaky(F.conv1d(low, sd[f"{name}.stack.main.{i}.weight"],
                             sd[f"{name}.stack.main.{i}.bias"], stride=2, padding=3))
        features.append(low)
        i += 1
    return features, F.conv1d(low, sd[f"{name}.judge.weight"], sd[f"{name}.judge.bias"], padding=1)


def filterbank_discriminator(x, feat, sd, bank, conditioning_channels=0):
    """discriminator/filterbank.py:163-202 -> ([8 maps, 3 maps, 3 maps], [3 judgements])"""
    a = F.conv1d(x.view(-1, 1, x.shape[-1]), bank, padding=bank.shape[-1] // 2)
    h = a
    if conditioning_channels > 0:
        h = torch.cat([h, F.interpolate(feat, size=h.shape[-1])], dim=1)
    full = []
    for i in range(len(FBD_MAIN)):
        h = leaky(F.conv1d(h, sd[f"main.{i}.weight"], sd[f"main.{i}.bias"], stride=2, padding=3))
        full.append(h)
    features = [full]
    judgements = [F.conv1d(h, sd["judge.weight"], sd["judge.bias"], padding=1)]
    for name in ("medium_res", "low_res"):
        f, j = _lowres_discriminator(a, feat, sd, name, conditioning_channels)
        features.append(f)
        judgements.append(j)
    return features, judgements


# ---------------------------------------------------------------------------------------
# GAN training step: featuresynth/train/train.py:26-42 (GeneratorTrainer.train), 63-74
# (DiscriminatorTrainer.train) with Adam(lr 1e-4, betas (0.5, 0.9)) from
# featuresynth/experiment/experiment.py:111-117, on the MelGanGenerator / MelGanDiscriminator
# pair (the discriminator ignores the conditioning features it is handed).
# ---------------------------------------------------------------------------------------
def adam_restated(params, grads, state, lr=1e-4, betas=(0.5, 0.9), eps=1e-8):
    """torch.optim.Adam.step (no amsgrad / weight decay) written out; state: dict name ->
    (step, exp_avg, exp_avg_sq), updated in place; returns the new parameter dict."""
    out = {}
    for name, p in params.items():
        g = grads[name]
        step, m, v = state.get(name, (0, torch.zeros_like(p), torch.zeros_like(p)))
        step += 1
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        bc1 = 1 - betas[0] ** step
        bc2 = 1 - betas[1] ** step
        denom = v.sqrt() / (bc2 ** 0.5) + eps
        out[name] = p - (lr / bc1) * (m / denom)
        state[name] = (step, m, v)
    return out


def _leaf(sd):
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}


def _melgan_gen(features, g_sd):
    return melgan_generator(features, g_sd)


def _melgan_disc(x, features, d_sd):
    return melgan_discriminator(x, d_sd)          # unconditioned: ignores the features


def _detach(x):
    return {k: v.detach() for k, v in x.items()} if isinstance(x, dict) else x.detach()


@torch.enable_grad()
def discriminator_train_step(g_sd, d_sd, samples, features, d_state, sub_loss=None,
                             gen_fn=_melgan_gen, disc_fn=_melgan_disc):
    """train/train.py:63-74 -> (d_loss, grads of D, new D state dict).  gen_fn / disc_fn select
    the model pair (default MelGanGenerator / MelGanDiscriminator)."""
    sub_loss = sub_loss or hinge_discriminator_loss
    d = _leaf(d_sd)
    with torch.no_grad():
        fake = gen_fn(features, g_sd)
    _, f_score = disc_fn(fake, features, d)
    _, r_score = disc_fn(samples, features, d)
    loss = mel_gan_disc_loss(r_score, f_score, gan_loss=sub_loss)
    names = list(d)
    grads = dict(zip(names, torch.autograd.grad(loss, [d[n] for n in names])))
    new = adam_restated({k: v.detach() for k, v in d.items()}, grads, d_state)
    return loss.item(), grads, new


@torch.enable_grad()
def generator_train_step(g_sd, d_sd, samples, features, g_state, sub_loss=None,
                         gen_fn=_melgan_gen, disc_fn=_melgan_disc):
    """train/train.py:26-42 -> (g_loss, fake, grads of G, new G state dict)"""
    sub_loss = sub_loss or hinge_generator_loss
    g = _leaf(g_sd)
    fake = gen_fn(features, g)
    f_features, f_score = disc_fn(fake, features, d_sd)
    with torch.no_grad():
        r_features, r_score = disc_fn(samples, features, d_sd)
    loss = mel_gan_gen_loss(r_features, f_features, r_score, f_score, gan_loss=sub_loss)
    names = list(g)
    grads = dict(zip(names, torch.autograd.grad(loss, [g[n] for n in names])))
    new = adam_restated({k: v.detach() for k, v in g.items()}, grads, g_state)
    return loss.item(), _detach(fake), grads, new


# ---------------------------------------------------------------------------------------
# Non-filterbank multiscale pair: MultiScaleGenerator (generator/multiscale.py:10-57, 180-251)
# and MultiScaleMultiResDiscriminator (discriminator/multiscale.py:10-67, 255-410)
# ---------------------------------------------------------------------------------------
MS_FACTORS = ((4, 4, 4, 4), (4, 4, 4, 2), (4, 4, 2, 2), (4, 2, 2, 2), (2, 2, 2, 2))
MS_G_CHANNELS = (512, 256, 128, 64, 32)
MS_D_CHANNELS = (1, 32, 64, 128, 256)


def _seeded(seed):
    import numpy as np
    rs = np.random.RandomState(seed)

    def w(*shape):
        return torch.from_numpy((rs.standard_normal(shape) * 0.02).astype(np.float32))

    def b(n):
        return torch.from_numpy((rs.standard_normal((n,)) * 0.01).astype(np.float32))
    return w, b


def multiscale_generator_state(seed, output_size, feature_channels=128):
    """state dict in the reference's key order (transposed_conv=True)"""
    w, b = _seeded(seed)
    sd = {"embedding.weight": w(512, feature_channels, 7), "embedding.bias": b(512)}
    for size, factors in zip(fb_band_sizes(output_size), MS_FACTORS):
        for i, s_ in enumerate(factors):
            cin, cout = MS_G_CHANNELS[i], MS_G_CHANNELS[i + 1]
            sd[f"channel_{size}.main.{2 * i}.conv.weight"] = w(cin, cout, 2 * s_)
            for j in range(3):
                sd[f"channel_{size}.main.{2 * i + 1}.main.{j}.weight"] = w(cout, cout, 3)
        sd[f"channel_{size}.to_samples.weight"] = w(1, 32, 7)
        sd[f"channel_{size}.to_samples.bias"] = b(1)
    return sd


def multiscale_generator(x, sd, output_size, recompose=False):
    """generator/multiscale.py:228-244 (forward), 53-56 (ChannelGenerator.forward),
    util/modules.py:120-137 (DilatedStack: x <- leaky(conv_d(x)[..., :L] + x)), 168-188."""
    T = x.shape[-1]
    e = leaky(F.conv1d(F.pad(x, (3, 3), mode="reflect"), sd["embedding.weight"],
                       sd["embedding.bias"]))
    results = {}
    for size, factors in zip(fb_band_sizes(output_size), MS_FACTORS):
        h = e
        for i, s_ in enumerate(factors):
            h = leaky(F.conv_transpose1d(h, sd[f"channel_{size}.main.{2 * i}.conv.weight"],
                                         stride=s_, padding=s_ // 2))
            for j, d in enumerate((1, 3, 9)):
                z = F.conv1d(h, sd[f"channel_{size}.main.{2 * i + 1}.main.{j}.weight"],
                             padding=d, dilation=d)[:, :, :h.shape[-1]]
                h = leaky(z + h)
        results[size] = F.conv1d(h, sd[f"channel_{size}.to_samples.weight"],
                                 sd[f"channel_{size}.to_samples.bias"], padding=3)
    if recompose:
        return fft_frequency_recompose(results, T * (output_size // T))
    return results


def multiscale_discriminator_state(seed, input_size, conditioning_channels=128, kernel_size=41,
                                   channel_judgements=True):
    w, b = _seeded(seed)
    sd = {}
    p = "multiscale."
    for size in fb_band_sizes(input_size):
        for i in range(4):
            sd[f"{p}channel_{size}.main.{i}.weight"] = w(MS_D_CHANNELS[i + 1], MS_D_CHANNELS[i], kernel_size)
            sd[f"{p}channel_{size}.main.{i}.bias"] = b(MS_D_CHANNELS[i + 1])
        if channel_judgements:
            for i in range(3):
                cin = 256 + conditioning_channels if i == 0 else 256
                sd[f"{p}channel_{size}.mj.{i}.weight"] = w(256, cin, 3)
                sd[f"{p}channel_{size}.mj.{i}.bias"] = b(256)
            sd[f"{p}channel_{size}.judge.weight"] = w(1, 256, 3)
            sd[f"{p}channel_{size}.judge.bias"] = b(1)
    for i in range(3):
        cin = 5 * 256 + conditioning_channels if i == 0 else 512
        sd[f"{p}final.{i}.weight"] = w(512, cin, 3)
        sd[f"{p}final.{i}.bias"] = b(512)
    sd[f"{p}judge.weight"] = w(1, 512, 3)
    sd[f"{p}judge.bias"] = b(1)
    return sd


def multiscale_multires_discriminator(x, feat, sd, input_size, decompose=True,
                                      channel_judgements=True, conditioning_channels=128,
                                      kernel_size=41):
    """discriminator/multiscale.py:50-67 (ChannelDiscriminator.forward), 327-375
    (MultiScaleDiscriminator.forward), 397-410 (flatten_multiscale_features=False)."""
    p = "multiscale."
    sizes = fb_band_sizes(input_size)
    bands = fft_frequency_decompose(x, min(sizes)) if decompose else x
    features, channels, judgements = [], [], []
    for size, factors in zip(sizes, MS_FACTORS):
        h = bands[size]
        f = []
        for i, s_ in enumerate(factors):
            h = leaky(F.conv1d(h, sd[f"{p}channel_{size}.main.{i}.weight"],
                               sd[f"{p}channel_{size}.main.{i}.bias"], stride=s_,
                               padding=kernel_size // 2))
            f.append(h)
        if channel_judgements:
            if conditioning_channels > 0:
                h = torch.cat([h, feat], dim=1)
            for i in range(3):
                h = leaky(F.conv1d(h, sd[f"{p}channel_{size}.mj.{i}.weight"],
                                   sd[f"{p}channel_{size}.mj.{i}.bias"], padding=1))
                f.append(h)
            judgements.append(F.conv1d(h, sd[f"{p}channel_{size}.judge.weight"],
                                       sd[f"{p}channel_{size}.judge.bias"], padding=1))
        features.append(f)
        channels.append(h)
    h = torch.cat(channels, dim=1)
    if conditioning_channels > 0:
        h = torch.cat([h, F.interpolate(feat, size=h.shape[-1])], dim=1)
    final = []
    for i in range(3):
        h = leaky(F.conv1d(h, sd[f"{p}final.{i}.weight"], sd[f"{p}final.{i}.bias"], padding=1))
        final.append(h)
    features.append(final)
    judgements.append(F.conv1d(h, sd[f"{p}judge.weight"], sd[f"{p}judge.bias"], padding=1))
    return features, judgements


# ---------------------------------------------------------------- data feed
def batch_stream(audio_chunks, spec_chunks, batch_size, feature_spec, anchor_feature, seed,
                 n_batches):
    """featuresynth/data/datastore.py:8-80 with the chunk features given as arrays (the
    reference reads them from its LMDB cache through `feature_funcs`): per example
    `random.choice` of a chunk, `random_slice` of the anchor feature, the aligned slice
    [start*ratio, end*ratio) of every other feature, zero padding, `conform` to
    (1, channels, size); batches are tuples in `feature_spec` order.  numpy, host only."""
    import random
    import numpy as np
    py_rng = random.Random(seed)
    np_rng = np.random.RandomState(seed)
    feats = {"audio": audio_chunks, "spectrogram": spec_chunks}

    def pad(x, length):
        if len(x) == length:
            return x
        widths = [(0, 0)] * x.ndim
        widths[0] = (0, length - len(x))
        return np.pad(x, widths, mode="constant")

    def conform(x, spec):
        size, channels = spec
        return x.T.reshape((-1, channels, size))

    anchor_size = feature_spec[anchor_feature][0]
    out = []
    for _ in range(n_batches):
        batch = {k: [] for k in feature_spec}
        for _ in range(batch_size):
            pick = py_rng.choice(range(len(audio_chunks)))
            arr = feats[anchor_feature][pick]
            room = len(arr) - anchor_size
            start = np_rng.randint(0, room) if room > 0 else 0
            end = start + anchor_size
            batch[anchor_feature].append(
                conform(pad(arr[start:end], anchor_size), feature_spec[anchor_feature]))
            for feat, spec in feature_spec.items():
                if feat == anchor_feature:
                    continue
                ratio = spec[0] // anchor_size
                ns, ne = start * ratio, end * ratio
                batch[feat].append(conform(pad(feats[feat][pick][ns:ne], ne - ns), spec))
        out.append(tuple(np.concatenate(batch[f], axis=0) for f in feature_spec))
    return out
