"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, imported through oracle/ref_harness.py) on seeded
synthetic inputs.  Run in the build container only:

    python -m oracle.make_golden

Inputs and weights are drawn from numpy RandomState (bit-stable across boxes) by
the same helpers the tests use, so the fixtures hold only seeds + reference
OUTPUTS.  Reference classes exercised (file:line):
  MelGanGenerator        featuresynth/generator/full.py:16-50
  ResidualStack          featuresynth/util/modules.py:391-405
  Audio2Mel              featuresynth/feature/feature.py:11-59
  MelGanDiscriminator    featuresynth/discriminator/melgan.py:7-27
  fft_frequency_*        featuresynth/audio/transform.py:50-115
  loss functions         featuresynth/loss/loss.py:5-79
  Generator/Discriminator trainers   featuresynth/train/train.py:26-74
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness, restate, bases, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    ref_harness.load()
    from featuresynth.generator.full import MelGanGenerator
    from featuresynth.util.modules import ResidualStack
    from featuresynth.feature.feature import Audio2Mel
    from featuresynth.discriminator.melgan import MelGanDiscriminator
    from featuresynth.audio.transform import (
        fft_frequency_decompose, fft_frequency_recompose)
    from featuresynth.loss import loss as ref_loss

    torch.set_grad_enabled(False)

    # ---- MelGanGenerator: small (B2,T8), biased variant, and cfg1 (B1,T64)
    for name, seed, B, T, biased in (("gen_b2_t8", 11, 2, 8, False),
                                     ("gen_b2_t8_bias", 12, 2, 8, True),
                                     ("gen_b3_t20_bias", 14, 3, 20, True),
                                     ("gen_cfg1_b1_t64", 0, 1, 64, False)):
        sd = restate.melgan_generator_state(seed)
        if biased:
            sd = restate.randomize_biases(sd, seed + 1000)
        g = MelGanGenerator(T, 128).eval()
        g.load_state_dict(sd)
        x = synth.mel_features(seed, B, T)
        y = g(x)
        save(name, seed=seed, B=B, T=T, biased=int(biased), y=y.numpy())

    # ---- ResidualStack alone
    for C, L, seed in ((32, 96, 21), (128, 64, 22)):
        rs = ResidualStack(C, [1, 3, 9]).eval()
        sd = synth.residual_stack_state(seed, C)
        rs.load_state_dict({k[len("s."):]: v for k, v in sd.items()})
        x = synth.randn(seed + 1, 2, C, L)
        save(f"resstack_c{C}", seed=seed, C=C, L=L, y=rs(x).numpy())

    # ---- Audio2Mel
    a2m = Audio2Mel(1024, 256, 1024, 22050, 128)
    save("mel_basis_22050_1024_128", mel_basis=a2m.mel_basis.numpy(),
         window=a2m.window.numpy())
    for name, seed, B, N in (("a2m_b2_n16384", 31, 2, 16384),
                             ("a2m_b3_n4000", 32, 3, 4000)):
        a = synth.uniform_audio(seed, B, N)
        save(name, seed=seed, B=B, N=N, y=a2m(a).numpy())

    # ---- MelGanDiscriminator (shared-weight, 3 scales)
    d = MelGanDiscriminator().eval()
    sd = restate.randomize_biases(restate.melgan_discriminator_state(41), 1041)
    d.load_state_dict(sd)
    x = synth.randn(42, 2, 1, 4096) * 0.1
    feats, judg = d(x)
    arrays = {"seed": 41, "N": 4096}
    for s, (fl, j) in enumerate(zip(feats, judg)):
        arrays[f"j{s}"] = j.numpy()
        for i, f in enumerate(fl):
            arrays[f"f{s}_{i}_shape"] = np.array(f.shape)
            arrays[f"f{s}_{i}_sub"] = f.numpy().reshape(-1)[::37]
            arrays[f"f{s}_{i}_abs_mean"] = f.abs().mean().item()
    save("disc_melgan_n4096", **arrays)

    # ---- losses (loss.py) on the discriminator outputs above vs a second input
    x2 = synth.randn(43, 2, 1, 4096) * 0.1
    feats2, judg2 = d(x2)
    save("losses_melgan",
         disc_hinge=ref_loss.mel_gan_disc_loss(judg, judg2).item(),
         disc_lsq=ref_loss.mel_gan_disc_loss(
             judg, judg2, gan_loss=ref_loss.least_squares_disc_loss).item(),
         feature=ref_loss.mel_gan_feature_loss(feats, feats2).item(),
         gen_hinge=ref_loss.mel_gan_gen_loss(feats, feats2, judg, judg2).item(),
         gen_lsq=ref_loss.mel_gan_gen_loss(
             feats, feats2, judg, judg2,
             gan_loss=ref_loss.least_squares_generator_loss).item())

    # ---- FFT band split / merge
    x = synth.randn(51, 2, 1, 8192) * 0.1
    bands = fft_frequency_decompose(x, 512)
    rec = fft_frequency_recompose(bands, 8192)
    arrays = {"seed": 51, "N": 8192, "min_size": 512, "recomposed": rec.numpy()}
    for k, v in bands.items():
        arrays[f"band_{k}"] = v.numpy()
    save("fft_bands_n8192", **arrays)

    # ---- Morlet filter bank analysis / synthesis (zounds restatement; the bank
    #      tensor itself is part of the fixture because its construction is unpinned)
    import zounds
    from featuresynth.generator.multiscale import FilterBankMultiScaleGenerator
    fbg = FilterBankMultiScaleGenerator(zounds.SR22050(), 128, 32, 8192,
                                        recompose=False)
    fb = fbg.channel_generators[8192].filter_bank
    x = synth.randn(61, 2, 1, 1024) * 0.1
    conv = fb.convolve(x)
    back = fb.transposed_convolve(conv)
    save("filterbank_n1024", seed=61, bank=fb.filter_bank.numpy(),
         conv_sub=conv.numpy().reshape(-1)[::29], conv_shape=np.array(conv.shape),
         back=back.numpy())


def fb_generator():
    """FilterBankMultiScaleGenerator, featuresynth/generator/multiscale.py:95-178."""
    ref_harness.load()
    import zounds
    from featuresynth.generator.multiscale import FilterBankMultiScaleGenerator
    torch.set_grad_enabled(False)
    for recompose in (False, True):
        g = FilterBankMultiScaleGenerator(zounds.SR22050(), 128, 8, 2048, recompose=recompose).eval()
        sd = restate.fb_generator_state(71, 2048)
        g.load_state_dict(sd)
        x = synth.mel_features(72, 2, 8)
        y = g(x)
        if recompose:
            save("fb_generator_recomposed_t8", seed=71, y=y.numpy())
        else:
            arrays = {"seed": 71, "sizes": np.array(list(y.keys()))}
            for k, v in y.items():
                arrays[f"band_{k}"] = v.numpy()
            for i, (size, cg) in enumerate(g.channel_generators.items()):
                arrays[f"bank_checksum_{size}"] = float(cg.filter_bank.filter_bank.double().abs().sum())
            save("fb_generator_t8", **arrays)


def fb_discriminator():
    """FilterBankMultiScaleDiscriminator, featuresynth/discriminator/multiscale.py:130-252."""
    ref_harness.load()
    import zounds
    from featuresynth.discriminator.multiscale import FilterBankMultiScaleDiscriminator
    torch.set_grad_enabled(False)
    d = FilterBankMultiScaleDiscriminator(2048, zounds.SR22050(), decompose=False,
                                          conditioning_channels=128).eval()
    sd = restate.fb_discriminator_state(81, 2048)
    d.load_state_dict(sd)
    bands = {s: synth.randn(82 + i, 2, 1, s) * 0.1 for i, s in enumerate(restate.fb_band_sizes(2048))}
    feat = synth.mel_features(90, 2, 8)
    feats, judg = d(bands, feat)
    arrays = {"seed": 81}
    for i, j in enumerate(judg):
        arrays[f"j{i}"] = j.numpy()
    for g, fl in enumerate(feats):
        for i, f in enumerate(fl):
            arrays[f"f{g}_{i}_shape"] = np.array(f.shape)
            arrays[f"f{g}_{i}_sub"] = f.numpy().reshape(-1)[::13]
    save("fb_discriminator_n2048", **arrays)


def filterbank_pair():
    """FilterBankGenerator (generator/filterbank.py:93-128) and FilterBankDiscriminator
    (discriminator/filterbank.py:114-202) as FilterBankExperiment / ConditionalFilterBankExperiment
    build them (experiment/filterbank.py:34-68, 93-106): unmodified reference classes, the
    zounds FilterBank stand-in of oracle/bases.py."""
    ref_harness.load()
    import zounds
    from featuresynth.generator.filterbank import FilterBankGenerator
    from featuresynth.discriminator.filterbank import FilterBankDiscriminator
    torch.set_grad_enabled(False)
    sr = zounds.SR22050()
    scale = zounds.LinearScale(zounds.FrequencyBand(20, sr.nyquist - 20), 128)
    fb = zounds.learn.FilterBank(sr, 511, scale, 0.9, normalize_filters=True, a_weighting=False)
    g = FilterBankGenerator(fb, 32, 8192, 128).eval()
    sd = restate.filterbank_generator_state(401)
    assert list(g.state_dict()) == list(sd), (list(g.state_dict()), list(sd))
    g.load_state_dict(sd)
    x = synth.mel_features(402, 2, 32)
    y = g(x)
    save("filterbank_generator_t32", seed=401, y=y.numpy(),
         bank_checksum=float(fb.filter_bank.double().abs().sum()),
         bank_sub=fb.filter_bank.numpy().reshape(-1)[::97])
    # ResidualStackFilterBankGenerator (generator/filterbank.py:8-90): the noise row comes from
    # torch.normal on the global host generator -- seeded right before the call
    from featuresynth.generator.filterbank import ResidualStackFilterBankGenerator
    rg = ResidualStackFilterBankGenerator(fb, 8, 2048, 128, add_weight_norm=True).eval()
    rsd = restate.resstack_filterbank_generator_state(411)
    assert list(rg.state_dict()) == list(rsd), (list(rg.state_dict())[:12], list(rsd)[:12])
    rg.load_state_dict(rsd)
    xr = synth.mel_features(412, 2, 8)
    torch.manual_seed(413)
    yr = rg(xr)
    torch.manual_seed(413)
    raw = torch.normal(0, 1, (1, 1, 2048))
    save("resstack_filterbank_generator_t8", seed=411, y=yr.numpy(), raw_noise=raw.numpy())
    for cond in (0, 128):
        d = FilterBankDiscriminator(fb, 8192, conditioning_channels=cond).eval()
        dsd = restate.filterbank_discriminator_state(403 + cond, conditioning_channels=cond)
        assert list(d.state_dict()) == list(dsd), (list(d.state_dict()), list(dsd))
        d.load_state_dict(dsd)
        a = synth.randn(404, 2, 1, 8192) * 0.1
        feat = synth.mel_features(405, 2, 32)
        feats, judg = d(a, feat)
        arrays = {"seed": 403 + cond}
        for i, j in enumerate(judg):
            arrays[f"j{i}"] = j.numpy()
        for gi, fl in enumerate(feats):
            for i, f in enumerate(fl):
                arrays[f"f{gi}_{i}_shape"] = np.array(f.shape)
                arrays[f"f{gi}_{i}_sub"] = f.numpy().reshape(-1)[::53]
        save("filterbank_discriminator_n8192" + ("_cond" if cond else ""), **arrays)


def realmelgan():
    """experiment/realmelgan.py Generator (48-89) and Discriminator (158-181)."""
    ref_harness.load()
    import featuresynth.experiment.realmelgan as rm
    torch.set_grad_enabled(False)
    g = rm.Generator(128, 32, n_residual_layers=3).eval()
    sd = restate.realmelgan_generator_state(101)
    assert list(g.state_dict()) == list(sd), "state-dict key order differs from the reference"
    g.load_state_dict(sd)
    x = synth.mel_features(102, 2, 8)
    y = g(x)
    save("realmelgan_gen_t8", seed=101, y=y.numpy(), n_keys=len(sd))
    d = rm.Discriminator(3, 16, 4, 4).eval()
    dsd = restate.realmelgan_discriminator_state(103)
    assert list(d.state_dict()) == list(dsd), "discriminator key order differs"
    d.load_state_dict(dsd)
    a = synth.randn(104, 2, 1, 4096) * 0.1
    feats, judg = d(a, None)
    arrays = {"seed": 103}
    for i, j in enumerate(judg):
        arrays[f"j{i}"] = j.numpy()
        for k, f in enumerate(feats[i]):
            arrays[f"f{i}_{k}_shape"] = np.array(f.shape)
            arrays[f"f{i}_{k}_sub"] = f.numpy().reshape(-1)[::41]
    save("realmelgan_disc_n4096", **arrays)
    # the pair's own losses (experiment/realmelgan.py:185-217) on D(real) vs D(second input)
    a2 = synth.randn(105, 2, 1, 4096) * 0.1
    feats2, judg2 = d(a2, None)
    save("realmelgan_losses",
         feature=float(rm.real_mel_gan_feature_loss(feats, feats2)),
         gen=float(rm.mel_gan_gen_loss(feats, feats2, judg, judg2)))


def train_step():
    """One DiscriminatorTrainer.train then one GeneratorTrainer.train (the `cycle` order of
    experiment/experiment.py:141-144) of the UNMODIFIED reference trainers
    (featuresynth/train/train.py:8-74) with torch Adam(1e-4, (0.5, 0.9)) on MelGanGenerator +
    MelGanDiscriminator (behind a 2-argument wrapper, the trainers pass the features too)."""
    ref_harness.load()
    from featuresynth.generator.full import MelGanGenerator
    from featuresynth.discriminator.melgan import MelGanDiscriminator
    from featuresynth.train.train import GeneratorTrainer, DiscriminatorTrainer
    from featuresynth.loss import loss as ref_loss

    class TwoArg(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.inner = MelGanDiscriminator()

        def forward(self, x, feat):
            return self.inner(x)

    B, T = 2, 8
    g = MelGanGenerator(T, 128)
    g.load_state_dict(restate.randomize_biases(restate.melgan_generator_state(111), 1111))
    d = TwoArg()
    d.inner.load_state_dict(restate.randomize_biases(restate.melgan_discriminator_state(112), 1112))
    g_optim = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = torch.optim.Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, ref_loss.mel_gan_disc_loss)
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, ref_loss.mel_gan_gen_loss)
    samples = synth.randn(113, B, 1, 256 * T) * 0.1
    features = synth.mel_features(114, B, T)
    arrays = {"B": B, "T": T}

    def sub(t):      # small tensors (biases) whole, large ones subsampled
        t = t.detach().reshape(-1)
        return (t if t.numel() <= 4096 else t[::257]).numpy().copy()

    g0 = {k: v.clone() for k, v in g.state_dict().items()}
    d0 = {k: v.clone() for k, v in d.inner.state_dict().items()}
    r = d_tr.train(samples, features)
    arrays["d_loss"] = r["d_loss"]
    for k, p in d.inner.named_parameters():
        arrays["dgrad." + k] = sub(p.grad)
        arrays["dgrad_norm." + k] = float(p.grad.norm())
        arrays["dnew." + k] = sub(p.detach() - d0[k])
    r = g_tr.train(samples, features)
    arrays["g_loss"] = r["g_loss"]
    arrays["fake"] = r["fake"][..., ::4]
    for k, p in g.named_parameters():
        arrays["ggrad." + k] = sub(p.grad)
        arrays["ggrad_norm." + k] = float(p.grad.norm())
        arrays["gnew." + k] = sub(p.detach() - g0[k])
    save("train_step_melgan_b2_t8", **arrays)


def multiscale():
    """MultiScaleGenerator (generator/multiscale.py:180-251) and MultiScaleMultiResDiscriminator
    (discriminator/multiscale.py:378-410) as wired by experiment/multiscale.py:81-93."""
    ref_harness.load()
    from featuresynth.generator.multiscale import MultiScaleGenerator
    from featuresynth.discriminator.multiscale import MultiScaleMultiResDiscriminator
    torch.set_grad_enabled(False)
    T, N = 8, 2048
    for recompose in (False, True):
        g = MultiScaleGenerator(128, T, N, transposed_conv=True, recompose=recompose).eval()
        sd = restate.multiscale_generator_state(171, N)
        assert list(g.state_dict()) == list(sd), "generator key order differs from the reference"
        g.load_state_dict(sd)
        y = g(synth.mel_features(172, 2, T))
        if recompose:
            save("ms_generator_recomposed_t8", seed=171, y=y.numpy())
        else:
            save("ms_generator_t8", seed=171, **{f"band_{k}": v.numpy() for k, v in y.items()})
    for name, kw in (("ms_discriminator_cond_n2048", dict(decompose=True, channel_judgements=True,
                                                          conditioning_channels=128)),
                     ("ms_discriminator_k9_n2048", dict(decompose=False, channel_judgements=True,
                                                        kernel_size=9))):
        d = MultiScaleMultiResDiscriminator(N, flatten_multiscale_features=False, **kw).eval()
        dsd = restate.multiscale_discriminator_state(
            173, N, conditioning_channels=kw.get("conditioning_channels", 0),
            kernel_size=kw.get("kernel_size", 41))
        assert list(d.state_dict()) == list(dsd), "discriminator key order differs"
        d.load_state_dict(dsd)
        if kw["decompose"]:
            x = synth.randn(174, 2, 1, N) * 0.1
        else:
            x = {s: synth.randn(175 + i, 2, 1, s) * 0.1 for i, s in enumerate(restate.fb_band_sizes(N))}
        feats, judg = d(x, synth.mel_features(180, 2, T))
        arrays = {"seed": 173, "n_groups": len(feats), "n_judgements": len(judg)}
        for i, j in enumerate(judg):
            arrays[f"j{i}"] = j.numpy()
        for gi, fl in enumerate(feats):
            arrays[f"n_f{gi}"] = len(fl)
            for i, f in enumerate(fl):
                arrays[f"f{gi}_{i}_shape"] = np.array(f.shape)
                arrays[f"f{gi}_{i}_sub"] = f.numpy().reshape(-1)[::13]
        save(name, **arrays)


def data_feed():
    """featuresynth/data/datastore.py:19-80 `batch_stream`, unmodified, over in-memory chunks:
    `iter_audio_chunks` is pointed at a chunk-id list and `feature_funcs` at array lookups (the
    reference's own extension points); python `random` and numpy's global generator are seeded."""
    import random
    ref_harness.load()
    import featuresynth.data.datastore as ds
    audio, spec = synth.feed_chunks(91)
    spec_spec = {"audio": (2048, 1), "spectrogram": (32, 16)}
    ds.iter_audio_chunks = lambda path, pattern: [(i, 0, 0) for i in range(len(audio))]
    funcs = {"audio": (lambda chunk: audio[chunk[0]], ()),
             "spectrogram": (lambda chunk: spec[chunk[0]], ())}
    random.seed(7)
    np.random.seed(7)
    stream = ds.batch_stream("unused", "unused", 6, spec_spec, "spectrogram", funcs)
    arrays = {"seed": 7, "chunk_seed": 91, "batch_size": 6}
    for i in range(3):
        a, s_ = next(stream)
        arrays[f"audio_{i}"] = a
        arrays[f"spectrogram_{i}"] = s_
    save("batch_stream", **arrays)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "data_feed":
        data_feed()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "multiscale":
        multiscale()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train_step":
        train_step()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "realmelgan":
        realmelgan()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "filterbank_pair":
        filterbank_pair()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fb_discriminator":
        fb_discriminator()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fb_generator":
        fb_generator()
        sys.exit(0)
    main()
