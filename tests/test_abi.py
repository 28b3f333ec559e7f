"""-m "not gpu": the C-ABI library loads and exports every symbol include/msb200.h
declares; host-side logic (descriptors, shapes, mel basis) without any compute call."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from music_synthesis_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    from music_synthesis_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "msb200.h")).read()
    declared = set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_status_strings_and_version(lib):
    assert lib.ms_version() >= 100
    assert lib.ms_strerror(0) == b"ok"
    assert b"workspace" in lib.ms_strerror(-3)
    assert lib.ms_launch_count() >= 0


def test_conv_geometry_host_logic(lib):
    from music_synthesis_b200 import ops
    # ResidualAtom convs keep the length; first conv on reflect-padded input
    assert ops.conv_out_len(ops.conv_desc(ops.MS_CONV, 1, 128, 128, 1000, 3, 9, 9)) == 1000
    assert ops.conv_out_len(ops.conv_desc(ops.MS_CONV, 1, 128, 512, 70, 7, 1, 0)) == 64
    # the reference's three upsampler geometries give exactly stride * L
    for k, s, p in ((16, 8, 4), (4, 2, 1), (8, 4, 2)):
        d = ops.conv_desc(ops.MS_CONVT, 1, 64, 32, 100, k, 1, p, s)
        assert ops.conv_out_len(d) == s * 100
    d = ops.conv_desc(ops.MS_CONV, 1, 256, 256, 64, 3, 1, 1)
    assert lib.ms_conv_packed_weight_bytes(ctypes.byref(d)) == 256 * 256 * 3 * 2
    assert lib.ms_audio2mel_frames(16384, 1024, 256) == 62
    assert lib.ms_audio2mel_frames(8192, 1024, 256) == 30
    assert lib.ms_audio2mel_frames(65536, 1024, 256) == 254
    assert lib.ms_melgan_packed_weight_bytes(128, 0) > 4_691_969 * 2 - 512
    assert lib.ms_melgan_workspace_bytes(2, 64, 128) == 2 * lib.ms_melgan_workspace_bytes(1, 64, 128)


def test_module_mirrors_keep_reference_state_dict_layout():
    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.feature.feature import Audio2Mel
    from oracle import restate
    g = MelGanGenerator(64, 128)
    ref = restate.melgan_generator_state(0)
    sd = g.state_dict()
    assert list(sd) == list(ref)            # same keys, same ORDER (pack order)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    assert sum(v.numel() for v in sd.values()) == 4_691_969
    a = Audio2Mel(1024, 256, 1024, 22050, 128)
    assert set(a.state_dict()) == {"mel_basis", "window"}
    assert tuple(a.mel_basis.shape) == (128, 513)


def test_product_mel_basis_equals_reference_fixture(golden):
    from music_synthesis_b200.feature.melbasis import mel_filterbank
    g = golden("mel_basis_22050_1024_128")
    assert np.array_equal(mel_filterbank(22050, 1024, 128, 0.0, None), g["mel_basis"])


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "music-synthesis_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_missing_library_fails_loudly(monkeypatch):
    from music_synthesis_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmsb200.so")
    with pytest.raises(ImportError):
        _lib.lib()


def test_packed_weight_layout_is_length_independent(lib):
    """NT/KB (the packed layout) must not change with the input length or batch."""
    from music_synthesis_b200 import ops
    for kind, cin, cout, k, s, p in ((ops.MS_CONV, 256, 256, 3, 1, 1), (ops.MS_CONVT, 512, 256, 16, 8, 4),
                                     (ops.MS_CONV, 128, 512, 7, 1, 0)):
        sizes = set()
        for lin in (8, 70, 129, 1024, 5000):
            d = ops.conv_desc(kind, 3, cin, cout, lin, k, 1, p, s)
            sizes.add(lib.ms_conv_packed_weight_bytes(ctypes.byref(d)))
        assert len(sizes) == 1 and 0 not in sizes


@pytest.mark.parametrize("k,s", [(7, 2), (7, 4), (41, 4), (41, 2), (9, 4)])
def test_strided_conv_weight_rewrite_is_exact_on_cpu(k, s):
    """host logic behind every strided conv on the tcgen05 path: Conv1d(k, stride s, pad k//2)
    == stride-1 conv over the space-to-depth input with the weight view of
    ms_strided_weight_view (its index map restated in tests/gpu_util.py and checked here with
    CPU torch ops; tests/test_gpu_kernels.py checks the kernel against the same restatement)"""
    import torch
    import torch.nn.functional as F
    from music_synthesis_b200 import ops
    from tests.gpu_util import strided_weight_view_ref
    torch.manual_seed(0)
    B, C, Co, L = 2, 3, 5, 37
    x = torch.randn(B, C, L)
    w = torch.randn(Co, C, k)
    ref = F.conv1d(x, w, stride=s, padding=k // 2)
    taps, pad = ops.strided_conv_geometry(k, s)
    w1 = strided_weight_view_ref(w, s)
    assert w1.shape == (Co, s * C, taps)
    lx = (L + s - 1) // s
    xp = F.pad(x, (0, lx * s - L))
    xs = xp.reshape(B, C, lx, s).permute(0, 3, 1, 2).reshape(B, s * C, lx)      # Y[i*C + c, u] = X[c, s*u + i]
    y = F.conv1d(xs, w1, padding=pad)[:, :, :lx]
    assert y.shape[-1] >= ref.shape[-1]
    assert torch.allclose(y[:, :, :ref.shape[-1]], ref, atol=1e-4)


def test_loss_scaler_policy(monkeypatch):
    import torch
    from music_synthesis_b200 import grad_ops
    from music_synthesis_b200.train.train import LossScaler
    monkeypatch.setattr(grad_ops, "NEEDS_LOSS_SCALE", True)
    sc = LossScaler(torch.device("cpu"), init_scale=1024.0, growth_interval=3)
    assert sc.enabled and float(sc.scale_dev) == 1024.0 and float(sc.inv_scale_dev) == 1 / 1024.0
    sc.update(True)                                   # overflow: halve, count the skipped step
    assert sc.scale == 512.0 and sc.skipped == 1 and float(sc.scale_dev) == 512.0
    for _ in range(3):
        sc.update(False)                              # growth_interval clean steps: double
    assert sc.scale == 1024.0
    monkeypatch.setattr(grad_ops, "NEEDS_LOSS_SCALE", False)
    off = LossScaler(torch.device("cpu"))
    off.update(True)
    assert not off.enabled and off.scale == 1.0       # bf16 backward: no scaling at all


def test_audio2mel_fft_passes_on_host(tmp_path):
    """csrc/a2m_fft.cuh (the register-resident 1024-point FFT of the Audio2Mel kernel) is
    __host__ __device__: tests/native/a2m_fft_host.cu runs the per-thread pass bodies on the CPU
    against a double-precision DFT (rel-L2 < 2e-6 or a non-zero exit status)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "a2m_fft_host")
    subprocess.run([nvcc, "-O2", "-Wno-deprecated-gpu-targets",
                    "-I", os.path.join(ROOT, "music-synthesis_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "native", "a2m_fft_host.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert float(r.stdout.split()[1]) < 2e-6


@pytest.mark.parametrize("batch,n,min_size", [(2, 65536, 4096), (2, 8192, 256), (3, 2048, 16),
                                              (1, 32768, 1024), (1, 16, 16), (1, 8, 4)])
@pytest.mark.parametrize("packed", [3, 2, 1, 0])
def test_fft_band_passes_on_host(tmp_path, batch, n, min_size, packed):
    """csrc/fft_passes.cuh (radix-2/4/16 Stockham passes with fused boundary loads / stores, the
    pass plan and the decompose / recompose sequences of fft_bands.cu, with real-input packing
    and without) is host-callable:
    tests/native/fft_bands_host.cu runs it on the CPU; compared with the oracle's restatement of
    featuresynth/audio/transform.py:50-115."""
    import shutil
    import subprocess
    import torch
    from oracle import restate
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = os.path.join(ROOT, "tests", "native", "_fft_bands_host")
    src = os.path.join(ROOT, "tests", "native", "fft_bands_host.cu")
    hdr = os.path.join(ROOT, "music-synthesis_b200", "csrc", "fft_passes.cuh")
    if (not os.path.exists(exe) or
            os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets",
                        "-I", os.path.dirname(hdr), "-o", exe, src], check=True)
    x = (np.random.RandomState(n + batch).randn(batch, n) * 0.1).astype(np.float32)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    x.tofile(fin)
    # packed: 3 = what the library runs (real-input packing, merge gathered by the inverse
    # transform's loader, two passes per launch), 2 = the same with one pass per launch,
    # 1 = packing with accumulation passes, 0 = full-length transforms
    r = subprocess.run([exe, fin, fout, str(batch), str(n), str(min_size), str(packed)],
                       stderr=subprocess.DEVNULL)
    assert r.returncode == 0
    out = np.fromfile(fout, dtype=np.float32)
    ref = restate.fft_frequency_decompose(torch.from_numpy(x)[:, None, :], min_size)
    o = 0
    for s in sorted(ref):
        got = out[o:o + batch * s].reshape(batch, s)
        o += batch * s
        want = ref[s][:, 0].numpy()
        assert np.linalg.norm(got - want) <= 3e-6 * np.linalg.norm(want), s
    want = restate.fft_frequency_recompose(ref, n)[:, 0].numpy()
    got = out[o:].reshape(batch, n)
    assert np.linalg.norm(got - want) <= 3e-6 * np.linalg.norm(want)


def test_convtranspose_column_order_is_a_bijection(tmp_path):
    """csrc/conv_gemm.cuh::convt_col (GEMM column of (phase, channel) in the polyphase
    ConvTranspose) and its two inverses, checked on the host for every stride / channel count
    the kernels accept: tests/native/convt_col_host.cu."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "convt_col_host")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets",
                    "-I", os.path.join(ROOT, "music-synthesis_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "native", "convt_col_host.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout)
