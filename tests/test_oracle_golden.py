"""Pins oracle/restate.py (the CPU restatement) to the reference's own outputs.

tests/golden/*.npz were produced by oracle/make_golden.py running the UNMODIFIED
reference modules; here the restatement is run on the same seeded inputs and must
agree to fp32 round-off.  (The reference has no golden vectors of its own,
SURVEY.md section 4, so these fixtures are the pin.)
"""
import numpy as np
import pytest
import torch

from oracle import restate, synth, bases



@pytest.fixture(autouse=True)
def _no_grad():
    """the pins run without a tape; scoped to this module's tests (a module-level
    torch.set_grad_enabled(False) leaked into every test collected after this file)"""
    with torch.no_grad():
        yield


def _sub(t):
    """the fixture keeps small tensors whole and every 257th element of large ones"""
    t = t.detach().reshape(-1)
    return t if t.numel() <= 4096 else t[::257]


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", ["gen_b2_t8", "gen_b2_t8_bias", "gen_b3_t20_bias",
                                  "gen_cfg1_b1_t64"])
def test_generator_matches_reference(golden, name):
    g = golden(name)
    seed, B, T = int(g["seed"]), int(g["B"]), int(g["T"])
    sd = restate.melgan_generator_state(seed)
    if int(g["biased"]):
        sd = restate.randomize_biases(sd, seed + 1000)
    y = restate.melgan_generator(synth.mel_features(seed, B, T), sd)
    assert y.shape == (B, 1, 256 * T)
    assert rel_l2(y, g["y"]) < 2e-6


@pytest.mark.parametrize("C", [32, 128])
def test_residual_stack_matches_reference(golden, C):
    g = golden(f"resstack_c{C}")
    seed, L = int(g["seed"]), int(g["L"])
    sd = synth.residual_stack_state(seed, C)
    y = restate.residual_stack(synth.randn(seed + 1, 2, C, L), sd, "s")
    assert rel_l2(y, g["y"]) < 1e-6


def test_mel_basis_restatement(golden):
    g = golden("mel_basis_22050_1024_128")
    mb = bases.librosa_mel(22050, 1024, 128, 0.0, None)
    assert mb.shape == (128, 513) and mb.dtype == np.float32
    assert np.array_equal(mb, g["mel_basis"])
    torchaudio = pytest.importorskip("torchaudio")
    tb = torchaudio.functional.melscale_fbanks(
        513, 0., 11025., 128, 22050, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(mb - tb).max() < 5e-7
    assert np.allclose(g["window"], torch.hann_window(1024).numpy(), atol=0)


@pytest.mark.parametrize("name", ["a2m_b2_n16384", "a2m_b3_n4000"])
def test_audio2mel_matches_reference(golden, name):
    g = golden(name)
    b = golden("mel_basis_22050_1024_128")
    a = synth.uniform_audio(int(g["seed"]), int(g["B"]), int(g["N"]))
    y = restate.audio2mel(a, torch.from_numpy(b["mel_basis"]),
                          torch.from_numpy(b["window"]))
    assert y.shape == g["y"].shape
    assert y.shape[-1] == (int(g["N"]) - 640) // 256 + 1
    assert np.abs(y.numpy() - g["y"]).max() < 2e-5


def test_discriminator_matches_reference(golden):
    g = golden("disc_melgan_n4096")
    sd = restate.randomize_biases(restate.melgan_discriminator_state(41), 1041)
    x = synth.randn(42, 2, 1, 4096) * 0.1
    feats, judg = restate.melgan_discriminator(x, sd)
    assert [j.shape[-1] for j in judg] == [16, 9, 5]
    for s, (fl, j) in enumerate(zip(feats, judg)):
        assert rel_l2(j, g[f"j{s}"]) < 2e-5
        for i, f in enumerate(fl):
            assert tuple(f.shape) == tuple(g[f"f{s}_{i}_shape"])
            assert rel_l2(f.reshape(-1)[::37], g[f"f{s}_{i}_sub"]) < 2e-5


def test_losses_match_reference(golden):
    g = golden("losses_melgan")
    sd = restate.randomize_biases(restate.melgan_discriminator_state(41), 1041)
    f1, j1 = restate.melgan_discriminator(synth.randn(42, 2, 1, 4096) * 0.1, sd)
    f2, j2 = restate.melgan_discriminator(synth.randn(43, 2, 1, 4096) * 0.1, sd)
    got = dict(
        disc_hinge=restate.mel_gan_disc_loss(j1, j2),
        disc_lsq=restate.mel_gan_disc_loss(j1, j2, restate.least_squares_disc_loss),
        feature=restate.mel_gan_feature_loss(f1, f2),
        gen_hinge=restate.mel_gan_gen_loss(f1, f2, j1, j2),
        gen_lsq=restate.mel_gan_gen_loss(f1, f2, j1, j2,
                                         restate.least_squares_generator_loss))
    for k, v in got.items():
        assert abs(float(v) - float(g[k])) <= 2e-5 * max(1.0, abs(float(g[k]))), k


def test_fft_bands_match_reference(golden):
    g = golden("fft_bands_n8192")
    x = synth.randn(51, 2, 1, 8192) * 0.1
    bands = restate.fft_frequency_decompose(x, 512)
    assert sorted(bands) == [512, 1024, 2048, 4096, 8192]
    for k, v in bands.items():
        assert rel_l2(v, g[f"band_{k}"]) < 2e-6
    rec = restate.fft_frequency_recompose(bands, 8192)
    assert rel_l2(rec, g["recomposed"]) < 2e-6
    # reference quirk: the split/merge is NOT perfectly reconstructing
    assert 1e-3 < rel_l2(rec, x) < 5e-2


def test_filterbank_matches_reference(golden):
    g = golden("filterbank_n1024")
    bank = torch.from_numpy(g["bank"])
    assert bank.shape == (128, 1, 128)
    # restated Morlet construction reproduces the bank the reference modules built
    scale = bases.LinearScale(bases.FrequencyBand(11025 / 2, 11025), 128)
    mine = bases.morlet_filter_bank(bases.SampleRate(22050), 128, scale, 0.05)
    assert np.array_equal(mine, g["bank"][:, 0, :])
    x = synth.randn(61, 2, 1, 1024) * 0.1
    conv = restate.filterbank_convolve(x, bank)
    assert tuple(conv.shape) == tuple(g["conv_shape"]) == (2, 128, 1025)
    assert rel_l2(conv.reshape(-1)[::29], g["conv_sub"]) < 2e-6
    back = restate.filterbank_transposed_convolve(conv, bank)
    assert rel_l2(back, g["back"]) < 2e-6


def test_reference_still_agrees_when_present():
    """In the build container, re-run the real reference live (not the fixture)."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("/root/reference not present on this box")
    ref_harness.load()
    from featuresynth.generator.full import MelGanGenerator
    sd = restate.melgan_generator_state(5)
    g = MelGanGenerator(4, 128).eval()
    g.load_state_dict(sd)
    x = synth.mel_features(6, 1, 4)
    assert rel_l2(restate.melgan_generator(x, sd), g(x)) < 2e-6


def test_fb_generator_matches_reference(golden):
    g = golden("fb_generator_t8")
    sd = restate.fb_generator_state(71, 2048)
    banks = restate.fb_banks()
    for size, bank in zip(restate.fb_band_sizes(2048), banks):
        assert abs(float(bank.double().abs().sum()) - float(g[f"bank_checksum_{size}"])) < 1e-6
    x = synth.mel_features(72, 2, 8)
    y = restate.filterbank_multiscale_generator(x, sd, banks, 2048)
    assert list(y) == list(g["sizes"]) == [2048, 1024, 512, 256, 128]
    for k, v in y.items():
        assert rel_l2(v, g[f"band_{k}"]) < 2e-6
    r = restate.filterbank_multiscale_generator(x, sd, banks, 2048, recompose=True)
    assert rel_l2(r, golden("fb_generator_recomposed_t8")["y"]) < 2e-6


def _fb_disc_inputs():
    bands = {s: synth.randn(82 + i, 2, 1, s) * 0.1
             for i, s in enumerate(restate.fb_band_sizes(2048))}
    return bands, synth.mel_features(90, 2, 8)


def test_fb_discriminator_matches_reference(golden):
    g = golden("fb_discriminator_n2048")
    sd = restate.fb_discriminator_state(81, 2048)
    bands, feat = _fb_disc_inputs()
    feats, judg = restate.filterbank_multiscale_discriminator(bands, feat, sd, restate.fb_banks(), 2048)
    assert len(judg) == 6 and [len(f) for f in feats] == [7, 7, 7, 7, 7, 3]
    for i, j in enumerate(judg):
        assert j.shape == (2, 1, 8) and rel_l2(j, g[f"j{i}"]) < 5e-6
    for gi, fl in enumerate(feats):
        for i, f in enumerate(fl):
            assert tuple(f.shape) == tuple(g[f"f{gi}_{i}_shape"])
            assert rel_l2(f.reshape(-1)[::13], g[f"f{gi}_{i}_sub"]) < 5e-6


def test_filterbank_experiment_pair_matches_reference(golden):
    """FilterBankGenerator / FilterBankDiscriminator: the restatement vs
    the unmodified reference classes over the 511-tap bank"""
    g = golden("filterbank_generator_t32")
    bank = restate.filterbank_experiment_bank()
    assert abs(float(bank.double().abs().sum()) - float(g["bank_checksum"])) < 1e-6
    assert np.abs(bank.numpy().reshape(-1)[::97] - g["bank_sub"]).max() < 1e-7
    sd = restate.filterbank_generator_state(401)
    y = restate.filterbank_generator(synth.mel_features(402, 2, 32), sd, bank)
    assert y.shape == (2, 1, 8192) and rel_l2(y, g["y"]) < 3e-6
    a = synth.randn(404, 2, 1, 8192) * 0.1
    feat = synth.mel_features(405, 2, 32)
    for cond, name in ((0, "filterbank_discriminator_n8192"),
                       (128, "filterbank_discriminator_n8192_cond")):
        gd = golden(name)
        dsd = restate.filterbank_discriminator_state(403 + cond, conditioning_channels=cond)
        feats, judg = restate.filterbank_discriminator(a, feat, dsd, bank, cond)
        assert [len(f) for f in feats] == [8, 3, 3]
        assert [tuple(j.shape) for j in judg] == [(2, 1, 32), (2, 1, 16), (2, 1, 4)]
        for i, j in enumerate(judg):
            assert rel_l2(j, gd[f"j{i}"]) < 2e-5
        for gi, fl in enumerate(feats):
            for i, f in enumerate(fl):
                assert tuple(f.shape) == tuple(gd[f"f{gi}_{i}_shape"])
                assert rel_l2(f.reshape(-1)[::53], gd[f"f{gi}_{i}_sub"]) < 2e-5


def test_resstack_filterbank_generator_matches_reference(golden):
    g = golden("resstack_filterbank_generator_t8")
    sd = restate.resstack_filterbank_generator_state(411)
    torch.manual_seed(413)
    raw = torch.normal(0, 1, (1, 1, 2048))
    assert np.array_equal(raw.numpy(), g["raw_noise"])      # the host generator is reproducible
    y = restate.resstack_filterbank_generator(synth.mel_features(412, 2, 8), sd,
                                              restate.filterbank_experiment_bank(), raw)
    assert y.shape == (2, 1, 2048) and rel_l2(y, g["y"]) < 5e-6


def test_product_511_tap_bank_equals_the_reference_bank(golden):
    """the product's own bank construction (audio/filterbank.py) for FilterBankExperiment"""
    from music_synthesis_b200.experiment.wirings import FilterBankExperiment
    g = golden("filterbank_generator_t32")
    fb = FilterBankExperiment.make_filter_bank()
    assert fb.kernel_size == 511 and fb.kp == 512 and fb.syn_nph == 32 and fb.syn_taps == 16 and fb.extra == 0
    assert abs(float(fb.filter_bank.double().abs().sum()) - float(g["bank_checksum"])) < 1e-4
    assert np.abs(fb.filter_bank.numpy().reshape(-1)[::97] - g["bank_sub"]).max() < 2e-7


def test_realmelgan_pair_matches_reference(golden):
    g = golden("realmelgan_gen_t8")
    sd = restate.realmelgan_generator_state(101)
    assert len(sd) == int(g["n_keys"]) == 126
    y = restate.realmelgan_generator(synth.mel_features(102, 2, 8), sd)
    assert y.shape == (2, 1, 2048) and rel_l2(y, g["y"]) < 3e-6
    gd = golden("realmelgan_disc_n4096")
    dsd = restate.realmelgan_discriminator_state(103)
    assert len(dsd) == 63
    feats, judg = restate.realmelgan_discriminator(synth.randn(104, 2, 1, 4096) * 0.1, dsd)
    assert [j.shape[-1] for j in judg] == [16, 8, 4]
    for i, j in enumerate(judg):
        assert rel_l2(j, gd[f"j{i}"]) < 2e-5
        assert len(feats[i]) == 6
        for k, f in enumerate(feats[i]):
            assert tuple(f.shape) == tuple(gd[f"f{i}_{k}_shape"])
            assert rel_l2(f.reshape(-1)[::41], gd[f"f{i}_{k}_sub"]) < 2e-5


def test_train_step_restatement_matches_reference_trainers(golden):
    """oracle D step then G step (restated trainers + restated Adam) vs the unmodified
    reference GeneratorTrainer / DiscriminatorTrainer + torch.optim.Adam"""
    g = golden("train_step_melgan_b2_t8")
    B, T = int(g["B"]), int(g["T"])
    g_sd = restate.randomize_biases(restate.melgan_generator_state(111), 1111)
    d_sd = restate.randomize_biases(restate.melgan_discriminator_state(112), 1112)
    samples = synth.randn(113, B, 1, 256 * T) * 0.1
    features = synth.mel_features(114, B, T)
    d_loss, d_grads, d_new = restate.discriminator_train_step(g_sd, d_sd, samples, features, {})
    assert abs(d_loss - float(g["d_loss"])) < 1e-5
    for k in d_sd:
        assert rel_l2(_sub(d_grads[k]), g["dgrad." + k]) < 1e-4, k
        assert rel_l2(_sub(d_new[k] - d_sd[k]), g["dnew." + k]) < 1e-3, k
    g_loss, fake, g_grads, g_new = restate.generator_train_step(g_sd, d_new, samples, features, {})
    assert abs(g_loss - float(g["g_loss"])) < 1e-5
    assert rel_l2(fake[..., ::4], g["fake"]) < 1e-5
    for k in g_sd:
        assert rel_l2(_sub(g_grads[k]), g["ggrad." + k]) < 1e-4, k
        assert rel_l2(_sub(g_new[k] - g_sd[k]), g["gnew." + k]) < 1e-3, k


def test_multiscale_pair_restatement_matches_reference(golden):
    T, N = 8, 2048
    sd = restate.multiscale_generator_state(171, N)
    x = synth.mel_features(172, 2, T)
    g = golden("ms_generator_t8")
    bands = restate.multiscale_generator(x, sd, N)
    for size, v in bands.items():
        assert rel_l2(v, g[f"band_{size}"]) < 5e-6
    assert rel_l2(restate.multiscale_generator(x, sd, N, recompose=True),
                  golden("ms_generator_recomposed_t8")["y"]) < 5e-6
    feat = synth.mel_features(180, 2, T)
    for name, kw in (("ms_discriminator_cond_n2048", dict(decompose=True, conditioning_channels=128)),
                     ("ms_discriminator_k9_n2048", dict(decompose=False, conditioning_channels=0,
                                                        kernel_size=9))):
        gd = golden(name)
        dsd = restate.multiscale_discriminator_state(173, N, kw["conditioning_channels"],
                                                     kw.get("kernel_size", 41))
        if kw["decompose"]:
            xin = synth.randn(174, 2, 1, N) * 0.1
        else:
            xin = {s: synth.randn(175 + i, 2, 1, s) * 0.1 for i, s in enumerate(restate.fb_band_sizes(N))}
        feats, judg = restate.multiscale_multires_discriminator(xin, feat, dsd, N, **kw)
        assert len(feats) == int(gd["n_groups"]) == 6 and len(judg) == int(gd["n_judgements"]) == 6
        for i, j in enumerate(judg):
            assert rel_l2(j, gd[f"j{i}"]) < 2e-5
        for gi, fl in enumerate(feats):
            assert len(fl) == int(gd[f"n_f{gi}"])
            for i, f in enumerate(fl):
                assert tuple(f.shape) == tuple(gd[f"f{gi}_{i}_shape"])
                assert rel_l2(f.reshape(-1)[::13], gd[f"f{gi}_{i}_sub"]) < 2e-5


def test_generator_gradient_tolerance_is_set_by_forward_rounding():
    """Why the training parity tests allow 5e-2 on generator gradients: the oracle with an EXACT
    fp32 backward but the tensor-core layers' forward operands rounded to fp16 (what the parity
    bar of 1e-3 on the waveform permits) already differs from the fp32 reference by ~4e-2 --
    LeakyReLU masks flip where activations are within the rounding error of zero.  The GPU
    path measures the same 4.2e-2 with bf16 AND with fp16 backward operands."""
    import torch.nn.functional as F
    B, T = 2, 8
    g_sd = restate.randomize_biases(restate.melgan_generator_state(121), 1121)
    d_sd = restate.randomize_biases(restate.melgan_discriminator_state(122), 1122)
    samples = synth.randn(123, B, 1, 256 * T) * 0.1
    features = synth.mel_features(125, B, T)
    kw = dict(sub_loss=restate.least_squares_generator_loss)
    _, _, ref, _ = restate.generator_train_step(g_sd, d_sd, samples, features, {}, **kw)
    c1, ct = F.conv1d, F.conv_transpose1d

    def r(t):                                   # round in the forward, identity in the backward
        return t + (t.half().float() - t).detach()

    def conv1d(x, w, b=None, *a, **k):
        if w.shape[0] == 1 or k.get("groups", 1) > 1 or x.shape[1] == 1:
            return c1(x, w, b, *a, **k)         # fp32 layers: mono convs, grouped / direct convs
        return c1(r(x), r(w), b, *a, **k)

    F.conv1d = conv1d
    F.conv_transpose1d = lambda x, w, b=None, *a, **k: ct(r(x), r(w), b, *a, **k)
    try:
        _, _, emu, _ = restate.generator_train_step(g_sd, d_sd, samples, features, {}, **kw)
    finally:
        F.conv1d, F.conv_transpose1d = c1, ct
    worst = max(rel_l2(emu[k], ref[k]) for k in ref)
    assert 5e-3 < worst < 8e-2, worst


def test_realmelgan_losses_match_reference(golden):
    g = golden("realmelgan_losses")
    dsd = restate.realmelgan_discriminator_state(103)
    f1, j1 = restate.realmelgan_discriminator(synth.randn(104, 2, 1, 4096) * 0.1, dsd)
    f2, j2 = restate.realmelgan_discriminator(synth.randn(105, 2, 1, 4096) * 0.1, dsd)
    feat = restate.real_mel_gan_feature_loss(f1, f2)
    gen = sum(restate.hinge_generator_loss(j) for j in j2) + 10 * feat
    assert abs(float(feat) - float(g["feature"])) < 1e-6 * max(1.0, abs(float(g["feature"])))
    assert abs(float(gen) - float(g["gen"])) < 1e-5 * max(1.0, abs(float(g["gen"])))


def test_reference_multiscale_representation_is_the_fft_bands_fixture(golden):
    """Build container only: the reference's `MultiScale.from_audio / to_audio`
    (audio/representation.py:82-103) reproduce the fft_bands_n8192 fixture exactly, so that
    fixture also pins the GPU `MultiScale` (tests/test_gpu_fft_bands.py)."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("/root/reference not present on this box")
    ref_harness.load()
    import os
    cwd = os.getcwd()
    from featuresynth.audio.representation import MultiScale
    os.chdir(cwd)
    g = golden("fft_bands_n8192")
    x = (synth.randn(51, 2, 1, 8192) * 0.1).numpy()
    ms = MultiScale.from_audio(x, 22050)
    assert sorted(ms.data) == [512, 1024, 2048, 4096, 8192]
    for k, v in ms.data.items():
        assert np.array_equal(v, g[f"band_{k}"])
    assert np.array_equal(ms.to_audio(), g["recomposed"].reshape(2, 8192))


FEED_SPEC = {"audio": (2048, 1), "spectrogram": (32, 16)}


def test_batch_stream_restatement_matches_reference(golden):
    """oracle/restate.py::batch_stream vs three batches of the unmodified reference's
    batch_stream (data/datastore.py:19-80) over the same chunks and seeds: bit-exact."""
    g = golden("batch_stream")
    audio, spec = synth.feed_chunks(int(g["chunk_seed"]))
    got = restate.batch_stream(audio, spec, int(g["batch_size"]), FEED_SPEC, "spectrogram",
                               int(g["seed"]), 3)
    for i, (a, s) in enumerate(got):
        assert np.array_equal(a, g[f"audio_{i}"]) and a.shape == (6, 1, 2048)
        assert np.array_equal(s, g[f"spectrogram_{i}"]) and s.shape == (6, 16, 32)
    # the fixture exercises the zero-padding path (chunks shorter than a crop)
    assert any((g[f"audio_{i}"][:, 0, -1] == 0).any() for i in range(3))


def test_data_feed_host_logic_matches_reference(golden):
    """The product's crop drawing and crop plans (data/datastore.py), with the gather itself
    emulated in numpy from the (origin, pitch, valid) rows ms_gather_crops consumes."""
    import random
    from music_synthesis_b200.data import DeviceAudioStore, draw_crops
    g = golden("batch_stream")
    audio, spec = synth.feed_chunks(int(g["chunk_seed"]))
    store = DeviceAudioStore(audio, spectrograms=spec, device="cpu")
    py_rng, np_rng = random.Random(int(g["seed"])), np.random.RandomState(int(g["seed"]))
    flat = {"audio": store.audio.numpy(), "spectrogram": store.spec.numpy()}
    for i in range(3):
        picks, starts = draw_crops(py_rng, np_rng, store.frames, 32, 6)
        plans = store.crop_plans(picks, starts, FEED_SPEC)
        for feat, (size, channels) in FEED_SPEC.items():
            out = np.zeros((6, channels, size), dtype=np.float32)
            for b, (origin, pitch, valid) in enumerate(plans[feat]):
                for c in range(channels):
                    lo = origin + c * pitch
                    out[b, c, :valid] = flat[feat][lo: lo + valid]
            assert np.array_equal(out, g[f"{feat}_{i}"]), (feat, i)
