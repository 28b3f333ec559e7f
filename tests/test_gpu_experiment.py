"""-m gpu: the experiment layer end to end on the GPU -- DeviceAudioStore -> batch_stream ->
preprocess_batch -> alternating trainer steps through `training_loop` -- with the first
discriminator and generator steps checked against the oracle's restated trainers on the very
batch the feed produced, and checkpoint / resume restoring the generator bit for bit."""
import numpy as np
import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_on():
    with torch.enable_grad():
        yield


def _store(seed=9):
    from music_synthesis_b200.data import DeviceAudioStore
    rs = np.random.RandomState(seed)
    chunks = [(rs.standard_normal(n) * 0.1).astype(np.float32) for n in (40000, 22050, 12000)]
    return DeviceAudioStore(chunks)


def _log_loss(exp, pre, result, iteration, elapsed):
    return {k: v for k, v in result.items() if isinstance(v, float)}


def test_melgan_experiment_training_loop_matches_oracle_steps(tmp_path, monkeypatch):
    from music_synthesis_b200.experiment import MultiScaleMelGanExperiment
    from music_synthesis_b200.train import training_loop
    torch.manual_seed(0)
    exp = MultiScaleMelGanExperiment().to("cuda")
    g0 = {k: v.detach().cpu().clone() for k, v in exp.generator.state_dict().items()}
    d0 = {k: v.detach().cpu().clone() for k, v in exp.discriminator.state_dict().items()}
    store = _store()
    batches = []

    def keep(exp_, pre, result, i, elapsed):
        batches.append(tuple(x.detach().cpu().clone() for x in pre))

    stream = exp.batch_stream(store, 4, seed=11)
    logs = []
    for i, elapsed, log in training_loop(stream, exp, "cuda", [_log_loss, keep]):
        logs.append(log)
        if i == 3:
            break
    assert [sorted(l) for l in logs] == [["d_loss"], ["g_loss"], ["d_loss"], ["g_loss"]]
    assert all(np.isfinite(list(l.values())[0]) for l in logs)
    assert batches[0][0].shape == (4, 1, 8192) and batches[0][1].shape == (4, 128, 32)
    # oracle: D step on batch 0 from the initial weights, then G step on batch 1
    d_loss, _, d1 = restate.discriminator_train_step(g0, d0, batches[0][0], batches[0][1], {})
    assert abs(logs[0]["d_loss"] - d_loss) < 2e-3 * abs(d_loss)
    g_loss, _, _, _ = restate.generator_train_step(g0, d1, batches[1][0], batches[1][1], {})
    assert abs(logs[1]["g_loss"] - g_loss) < 2e-3 * max(1.0, abs(g_loss))
    moved = [float((v.detach().cpu() - g0[k]).abs().max())
             for k, v in exp.generator.state_dict().items()]
    assert max(moved) > 5e-5                       # two Adam steps of lr 1e-4
    # checkpoint / resume (reference file names, cwd-relative)
    monkeypatch.chdir(tmp_path)
    exp.checkpoint("t_")
    feats = batches[0][1].cuda()
    with torch.no_grad():
        y0 = exp.generator(feats).clone()
    next(exp.training_steps)(batches[2][0].cuda(), feats)      # D
    next(exp.training_steps)(batches[2][0].cuda(), feats)      # G: weights move
    with torch.no_grad():
        assert not torch.equal(exp.generator(feats), y0)
    exp.resume("t_")
    with torch.no_grad():
        assert torch.equal(exp.generator(feats), y0)


def test_filterbank_multiscale_experiment_runs_on_band_dictionaries():
    """cfg5 wiring: `preprocess_batch` turns the feed's audio into MultiScale bands on the GPU
    (no host round trip), the trainers consume the dictionaries, 'fake' comes back per band."""
    from music_synthesis_b200.experiment import FilterBankMultiscaleExperiment
    from music_synthesis_b200.train import training_loop
    torch.manual_seed(0)
    exp = FilterBankMultiscaleExperiment().to("cuda")
    stream = exp.batch_stream(_store(), 2, seed=5)
    seen = []

    def keep(exp_, pre, result, i, elapsed):
        seen.append((pre, result))

    for i, elapsed, log in training_loop(stream, exp, "cuda", [_log_loss, keep]):
        assert np.isfinite(list(log.values())[0])
        if i == 1:
            break
    bands, feats = seen[0][0]
    assert sorted(bands) == [512, 1024, 2048, 4096, 8192] and bands[512].is_cuda
    fake = seen[1][1]["fake"]
    assert sorted(fake) == sorted(bands) and fake[8192].shape == (2, 1, 8192)
    audio = exp.audio_representation(fake, exp.samplerate).to_audio()
    assert audio.shape == (2, 8192) and np.isfinite(audio).all()
