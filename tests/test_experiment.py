"""-m "not gpu": the experiment layer (featuresynth/experiment/*.py) is host glue, so its
drop-in contract is checked on the CPU: constructor wiring, initialiser, specs, checkpoint file
names and -- in the build container, against the UNMODIFIED reference experiments imported
through oracle/ref_harness.py -- state-dict layouts and checkpoint files exchanged both ways."""
import os

import numpy as np
import pytest
import torch

EXPERIMENTS = [
    ("melgan", "MultiScaleMelGanExperiment"),
    ("realmelgan", "RealMelGanExperiment"),
    ("multiscale", "FilterBankMultiscaleExperiment"),
    ("multiscale", "MultiScaleNoDeRecompose"),
    ("multiscale", "MultiScaleNoDeRecomposeUnconditionedShortKernel"),
    ("filterbank", "FilterBankExperiment"),
    ("filterbank", "ConditionalFilterBankExperiment"),
    ("filterbank", "AlternateFilterBankExperiment"),
]


def _ours(name):
    import music_synthesis_b200.experiment as ex
    return getattr(ex, name)


def test_experiment_contract_on_host():
    exp = _ours("RealMelGanExperiment")()
    assert exp.feature_spec == {"audio": (8192, 1), "spectrogram": (32, 128)}
    assert exp.inference_spec == {"audio": (32768, 1), "spectrogram": (128, 128)}
    assert exp._gen_name("a_") == "trained_models/a_realmelgan_gen.dat"
    assert exp._disc_name() == "trained_models/realmelgan_disc.dat"
    # weights_init ran over both networks (experiment/init.py:3-9): N(0, 0.02), zero bias
    first = exp.discriminator.model["disc_0"].model["layer_0"][1]
    assert float(first.bias.abs().max()) == 0.0
    w = torch.cat([p.detach().reshape(-1) for n, p in exp.generator.named_parameters()
                   if n.endswith("weight_v")])
    assert abs(float(w.std()) - 0.02) < 1e-3 and abs(float(w.mean())) < 1e-3
    # the reference refuses to build without feature funcs; so does the mirror
    from music_synthesis_b200.experiment import Experiment
    with pytest.raises(ValueError):
        Experiment(exp.generator, exp.discriminator, 1e-4, 32, None, None, None)
    # trainers need HBM-resident parameters: on a box without a GPU this must fail loudly
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            exp.generator_trainer


def test_raw_audio_representation_and_batch_preprocessing():
    exp = _ours("MultiScaleMelGanExperiment")()
    samples = np.zeros((3, 1, 8192), dtype=np.float32)
    feats = np.zeros((3, 128, 32), dtype=np.float32)
    s, f = exp.preprocess_batch((samples, feats))
    assert s is samples and f is feats
    assert exp.audio_representation(samples, 22050).to_audio().shape == (3, 8192)


@pytest.mark.parametrize("module,name", EXPERIMENTS)
def test_checkpoints_are_exchangeable_with_the_reference(tmp_path, monkeypatch, module, name):
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("/root/reference not present on this box")
    ref_harness.load()
    import importlib
    ref_cls = getattr(importlib.import_module("featuresynth.experiment." + module), name)
    monkeypatch.chdir(tmp_path)
    os.makedirs("trained_models")
    ref, ours = ref_cls(), _ours(name)()
    assert ours._gen_name("p_") == ref._gen_name("p_")
    assert ours._disc_name("p_") == ref._disc_name("p_")
    assert ours.feature_spec == ref.feature_spec and ours.inference_spec == ref.inference_spec
    for net in ("generator", "discriminator"):
        a, b = getattr(ref, net).state_dict(), getattr(ours, net).state_dict()
        assert list(a) == list(b), net
        assert [tuple(v.shape) for v in a.values()] == [tuple(v.shape) for v in b.values()]
    # reference -> ours
    ref.checkpoint("r_")
    ours.CHECKPOINT_DIR = "trained_models"
    ours.resume("r_")
    for k, v in ref.generator.state_dict().items():
        assert torch.equal(v, ours.generator.state_dict()[k])
    # ours -> reference (fresh weights first, so the load is observable)
    ours.generator.apply(ours.generator_init)
    ours.checkpoint("o_")
    ref.resume("o_")
    for k, v in ours.discriminator.state_dict().items():
        assert torch.equal(v, ref.discriminator.state_dict()[k])
    for k, v in ours.generator.state_dict().items():
        assert torch.equal(v, ref.generator.state_dict()[k])
