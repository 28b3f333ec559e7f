"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import harness for the *unmodified* reference at /root/reference (read-only).

The reference (JohnVinyard/music-synthesis) is 2020-era research code: it imports
zounds / librosa / lws / lmdb / soundfile / boto3 / deploygraph / matplotlib (none
installed here, no network) and calls torch APIs that were removed since
(`torch.rfft/irfft`, real-valued `torch.stft`, `scipy.signal.hann`).  None of the
hot-path `nn.Module`s need those packages to *compute*, so this harness registers
permissive stub modules, three torch/scipy shims, and numpy restatements of the
two third-party fixed bases (`librosa.filters.mel`, `zounds.learn.FilterBank`),
then imports `featuresynth` from /root/reference as-is.  No reference file is
edited or copied.

It only works in the build container (where /root/reference exists).  It is used
by `oracle/make_golden.py` to generate the fixtures under `tests/golden/`, and by
`-m "not gpu"` tests (skipped when /root/reference is absent) to pin
`oracle/restate.py` against the real reference.

Reference call sites the shims serve:
  featuresynth/feature/feature.py:7,27-29   librosa.filters.mel (positional)
  featuresynth/feature/feature.py:47-55     legacy real-valued torch.stft
  featuresynth/audio/transform.py:51,70,87,99  torch.rfft / torch.irfft
  featuresynth/generator/ddsp.py:62         scipy.signal.hann
  featuresynth/generator/multiscale.py:151-164  zounds FilterBank construction
"""
import os
import sys
import types
import tempfile

REFERENCE_ROOT = "/root/reference"

_STUBS = [
    "zounds", "zounds.learn", "zounds.spectral", "zounds.spectral.functional",
    "zounds.nputil", "librosa", "librosa.filters", "librosa.util", "lws", "lmdb",
    "soundfile", "deploygraph", "deploygraph.aws", "botocore", "botocore.config",
    "boto3", "matplotlib", "matplotlib.pyplot",
]


class _Dummy:
    """Permissive placeholder: callable, subscriptable, attribute-able."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __getitem__(self, k):
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "featuresynth"))


_loaded = None


def load():
    """Return the imported reference `featuresynth` package (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    import numpy as np
    import scipy.signal
    import scipy.signal.windows
    import torch

    from oracle import bases

    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _StubModule(name)
    for name in _STUBS:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])

    # --- real replacements inside the stubs --------------------------------
    sys.modules["librosa.filters"].mel = bases.librosa_mel
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    z = sys.modules["zounds"]
    z.SR22050 = lambda: bases.SampleRate(22050)
    z.SR11025 = lambda: bases.SampleRate(11025)
    z.SampleRate = bases.SampleRate
    z.FrequencyBand = bases.FrequencyBand
    z.LinearScale = bases.LinearScale
    sys.modules["zounds.learn"].FilterBank = bases.FilterBank
    z.learn = sys.modules["zounds.learn"]

    # --- API-rot shims ---------------------------------------------------------
    if not hasattr(scipy.signal, "hann"):
        scipy.signal.hann = scipy.signal.windows.hann

    if not getattr(torch.stft, "_oracle_shim", False):
        _real_stft = torch.stft

        def _stft(*a, **k):
            if k.get("return_complex") is None:
                k["return_complex"] = True
                return torch.view_as_real(_real_stft(*a, **k))
            return _real_stft(*a, **k)

        _stft._oracle_shim = True
        torch.stft = _stft

    if not hasattr(torch, "rfft"):
        def _rfft(input, signal_ndim, normalized=False, onesided=True):
            x = input
            assert signal_ndim == 1 and onesided
            return torch.view_as_real(
                torch.fft.rfft(x, norm="ortho" if normalized else "backward"))

        def _irfft(input, signal_ndim, normalized=False, onesided=True,
                   signal_sizes=None):
            x = input
            assert signal_ndim == 1 and onesided
            n = signal_sizes[0] if signal_sizes is not None else None
            return torch.fft.irfft(
                torch.view_as_complex(x.contiguous()), n=n,
                norm="ortho" if normalized else "backward")

        torch.rfft = _rfft
        torch.irfft = _irfft

    # --- import ----------------------------------------------------------------
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # experiment/__init__.py pulls in report.py (deploygraph.Requirement subclass,
    # AWS publishing): skip it but keep the sub-modules importable.
    exp = types.ModuleType("featuresynth.experiment")
    exp.__path__ = [os.path.join(REFERENCE_ROOT, "featuresynth", "experiment")]
    sys.modules["featuresynth.experiment"] = exp

    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="oracle_ref_")
    os.chdir(scratch)  # feature/feature.py:62 opens an LMDB dir in cwd
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import featuresynth  # noqa: F401
            import featuresynth.generator.full  # noqa: F401
            import featuresynth.discriminator.melgan  # noqa: F401
            import featuresynth.feature.feature  # noqa: F401
            import featuresynth.audio.transform  # noqa: F401
            import featuresynth.loss.loss  # noqa: F401
            import featuresynth.train.train  # noqa: F401
            import featuresynth.experiment.init  # noqa: F401
    finally:
        os.chdir(cwd)
    _loaded = sys.modules["featuresynth"]
    return _loaded
