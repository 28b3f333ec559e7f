"""-m gpu: backward-pass primitives of the training step against
CPU autograd on the same seeded inputs.  16-bit GEMM operands: the reference gradient is
computed in fp32 from fp32 inputs, tolerances state the operand rounding."""
import pytest
import torch
import torch.nn.functional as F

from oracle import synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_on():
    # other test modules switch autograd off process-wide
    with torch.enable_grad():
        yield


def _r16(x, fmt):
    return x.to(torch.bfloat16 if fmt == 1 else torch.float16).float()


@pytest.mark.parametrize("fmt_dz,fmt_x", [(0, 0), (1, 1)])
def test_wgrad_operand_formats(fmt_dz, fmt_x):
    """MN-major tcgen05 weight gradient, fp16 and bf16 operands (the hardware faults on a mix),
    vs an exact fp64 sum of the rounded operands"""
    from music_synthesis_b200 import ops, grad_ops
    B, Co, Ci, L, K, d = 3, 128, 64, 300, 3, 3
    dz = synth.randn(1, B, Co, L)
    x = synth.randn(2, B, Ci, L)
    dz16 = ops.pack_ncl(dz.cuda(), operand=fmt_dz)
    x16 = ops.pack_ncl(x.cuda(), operand=fmt_x)
    dw = grad_ops.conv_wgrad(dz16, x16, (Co, Ci, K), dilation=d, pad=d, fmt_dz=fmt_dz, fmt_x=fmt_x)
    xr = F.pad(_r16(x, fmt_x).double(), (d, d))
    dzr = _r16(dz, fmt_dz).double()
    ref = torch.stack([torch.einsum("bol,bil->oi", dzr, xr[:, :, k * d:k * d + L]) for k in range(K)], 2)
    err = rel_l2(dw, ref.float())
    print("wgrad", fmt_dz, fmt_x, err)
    assert err < 1e-5


@pytest.mark.parametrize("B,Co,Ci,L,K,d,pad", [(2, 32, 32, 700, 3, 9, 9), (2, 256, 256, 130, 3, 1, 1),
                                               (1, 512, 128, 40, 7, 1, 0), (2, 1024, 1024, 33, 5, 1, 2),
                                               (5, 64, 64, 128, 3, 1, 1),
                                               # short clips: several per tile / K chunk (folded)
                                               (5, 64, 64, 20, 3, 3, 3), (7, 128, 256, 9, 5, 1, 2),
                                               (33, 1024, 1024, 5, 5, 1, 2), (4, 32, 32, 61, 3, 1, 1)])
def test_conv_backward_matches_autograd(B, Co, Ci, L, K, d, pad):
    from music_synthesis_b200 import ops, grad_ops
    x = synth.randn(3, B, Ci, L).requires_grad_()
    w = (synth.randn(4, Co, Ci, K) * 0.05).requires_grad_()
    y = F.conv1d(x, w, None, 1, pad, d)
    dy = synth.randn(5, *y.shape) * 1e-4          # small gradients: must survive the 16-bit operand
    y.backward(dy)
    dz16 = ops.pack_ncl(dy.cuda(), operand=grad_ops.GRAD_FMT)
    x16 = ops.pack_ncl(x.detach().cuda())
    dw = grad_ops.conv_wgrad(dz16, x16, tuple(w.shape), dilation=d, pad=pad)
    dx32 = grad_ops.conv_dgrad(w.detach().cuda(), dz16, ops.MS_CONV, dilation=d, pad=pad)
    e_w, e_x = rel_l2(dw, w.grad), rel_l2(ops.unpack_blk32(dx32), x.grad)
    print("conv bwd", (B, Co, Ci, L, K, d, pad), e_w, e_x)
    assert e_w < 5e-3 and e_x < 5e-3


@pytest.mark.parametrize("B,Ci,Co,L,s,p", [(2, 512, 256, 32, 8, 4), (3, 128, 64, 200, 2, 1),
                                           (2, 64, 32, 300, 2, 1), (2, 128, 128, 64, 4, 2)])
def test_conv_transpose_backward_matches_autograd(B, Ci, Co, L, s, p):
    from music_synthesis_b200 import ops, grad_ops
    x = synth.randn(6, B, Ci, L).requires_grad_()
    w = (synth.randn(7, Ci, Co, 2 * s) * 0.05).requires_grad_()
    y = F.conv_transpose1d(x, w, None, s, p)
    dy = synth.randn(8, *y.shape) * 1e-4
    y.backward(dy)
    dy32 = grad_ops.pack_ncl32(dy.cuda())
    dzs16, db = grad_ops.act_bwd(dy32, s2d=s)
    assert rel_l2(db, dy.sum((0, 2))) < 1e-4
    x16 = ops.pack_ncl(x.detach().cuda())
    dw = grad_ops.convt_wgrad(x16, dzs16, tuple(w.shape), s, p)
    dx32 = grad_ops.conv_dgrad(w.detach().cuda(), dzs16, ops.MS_CONVT, stride=s, pad=p)
    e_w, e_x = rel_l2(dw, w.grad), rel_l2(ops.unpack_blk32(dx32), x.grad)
    print("convT bwd", (B, Ci, Co, L, s, p), e_w, e_x)
    assert e_w < 5e-3 and e_x < 5e-3


def test_act_bwd_masks_and_bias():
    from music_synthesis_b200 import ops, grad_ops
    B, C, L = 2, 64, 1500
    z = synth.randn(9, B, C, L).requires_grad_()
    y = F.leaky_relu(z, 0.2)
    dy = synth.randn(10, B, C, L)
    y.backward(dy)
    dy32 = grad_ops.pack_ncl32(dy.cuda())
    y16 = ops.pack_ncl(y.detach().cuda())
    dz16, db = grad_ops.act_bwd(dy32, sign16=y16, fmt=ops.MS_BF16)
    assert rel_l2(ops.unpack_blk16(dz16, ops.MS_BF16), z.grad) < 4e-3
    assert rel_l2(db, z.grad.sum((0, 2))) < 1e-5
    # residual-branch form: sign of (y - x)
    x = synth.randn(11, B, C, L)
    ya = grad_ops.pack_ncl32((x + y.detach()).cuda())
    xb = grad_ops.pack_ncl32(x.cuda())
    dz16b, _ = grad_ops.act_bwd(dy32, ya32=ya, yb32=xb, fmt=ops.MS_F16, want_bias=False)
    assert rel_l2(ops.unpack_blk16(dz16b), z.grad) < 2e-3


@pytest.mark.parametrize("cin,cout,k,s,pad,g,L", [(1, 16, 15, 1, 7, 1, 1000), (16, 64, 41, 4, 20, 4, 1000),
                                                  (64, 256, 41, 4, 20, 16, 250), (1024, 1024, 41, 4, 20, 256, 37)])
def test_direct_conv_backward(cin, cout, k, s, pad, g, L):
    from music_synthesis_b200 import ops, grad_ops
    B = 3
    x = synth.randn(12, B, cin, L).requires_grad_()
    w = (synth.randn(13, cout, cin // g, k) * 0.1).requires_grad_()
    b = (synth.randn(14, cout) * 0.1).requires_grad_()
    y = F.leaky_relu(F.conv1d(x, w, b, s, pad, 1, g), 0.2)
    dy = synth.randn(15, *y.shape)
    y.backward(dy)
    yg = ops.conv1d_direct(x.detach().cuda(), w.detach().cuda(), b.detach().cuda(), s, pad, g, leaky=True)
    assert rel_l2(yg, y.detach()) < 1e-5
    dx, dw, db = grad_ops.conv1d_direct_bwd(dy.cuda(), yg, x.detach().cuda(), w.detach().cuda(), s, pad, g, True)
    assert rel_l2(dx, x.grad) < 1e-5
    assert rel_l2(dw, w.grad) < 1e-5
    assert rel_l2(db, b.grad) < 1e-5


@pytest.mark.parametrize("cin,k,pad,tanh", [(32, 7, 3, True), (1024, 3, 1, False)])
def test_conv_to_mono_backward(cin, k, pad, tanh):
    from music_synthesis_b200 import ops, grad_ops
    B, L = 2, 777
    x = synth.randn(16, B, cin, L).requires_grad_()
    w = (synth.randn(17, 1, cin, k) * 0.05).requires_grad_()
    b = (synth.randn(18, 1) * 0.1).requires_grad_()
    y = F.conv1d(x, w, b, 1, pad)
    if tanh:
        y = torch.tanh(y)
    dy = synth.randn(19, *y.shape)
    y.backward(dy)
    x32 = grad_ops.pack_ncl32(x.detach().cuda())
    yg = ops.conv_to_mono(x32, w.detach().cuda(), b.detach().cuda(), k, pad, tanh)
    dx32, dw, db = grad_ops.conv_to_mono_bwd(dy.cuda(), yg if tanh else None, x32, w.detach().cuda(), k, pad)
    assert rel_l2(ops.unpack_blk32(dx32), x.grad) < 1e-5
    assert rel_l2(dw, w.grad) < 1e-5
    assert rel_l2(db, b.grad) < 1e-5


@pytest.mark.parametrize("include_pad,pad", [(True, 2), (False, 1)])
def test_avg_pool_backward(include_pad, pad):
    from music_synthesis_b200 import grad_ops
    x = synth.randn(20, 2, 3, 1001).requires_grad_()
    y = F.avg_pool1d(x, 4, 2, pad, count_include_pad=include_pad)
    dy = synth.randn(21, *y.shape)
    y.backward(dy)
    dx = grad_ops.avg_pool1d_bwd(dy.cuda(), 1001, 4, 2, pad, include_pad)
    assert rel_l2(dx, x.grad) < 1e-6


def test_loss_gradients_and_adam():
    from music_synthesis_b200 import grad_ops
    a = synth.randn(22, 4, 1, 500).requires_grad_()
    b = synth.randn(23, 4, 1, 500).requires_grad_()
    g = torch.tensor([0.7])
    cases = {0: lambda: F.l1_loss(a, b), 1: lambda: (F.relu(1 - a) + F.relu(1 + b)).mean(),
             2: lambda: (-a).mean(), 3: lambda: 0.5 * (((a - 1) ** 2).mean() + (b ** 2).mean()),
             4: lambda: 0.5 * ((a - 1) ** 2).mean()}
    for mode, fn in cases.items():
        a.grad = b.grad = None
        (fn() * 3.0).backward(g[0])
        da, db = grad_ops.reduce_bwd(mode, a.detach().cuda(), b.detach().cuda(), 3.0, g.cuda(), True, True)
        assert rel_l2(da, a.grad) < 1e-6, mode
        if b.grad is not None:
            assert rel_l2(db, b.grad) < 1e-6, mode
    p = synth.randn(24, 10000) * 0.02
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.5, 0.9))
    pg = p.clone().cuda()
    m, v = torch.zeros_like(pg), torch.zeros_like(pg)
    for step in range(1, 4):
        grad = synth.randn(24 + step, 10000) * 1e-3
        ref.grad = grad.clone()
        opt.step()
        grad_ops.adam_step(pg, (grad * 2).cuda(), m, v, 1e-4, 0.5, 0.9, 1e-8, step, grad_scale=0.5)
    assert rel_l2(pg - p.cuda(), ref.detach() - p) < 1e-5


@pytest.mark.parametrize("L", [3, 5, 8, 13])
def test_short_sequences_top_of_discriminator(L):
    """the top of the discriminator at its coarse scales runs on 3..8 time steps
    (discriminator/melgan.py:20-25: 8192 -> 4097 -> 2049 samples, /256): dense k5 layer + judge,
    forward and backward, vs CPU autograd"""
    from music_synthesis_b200 import autograd as ag
    B, C = 2, 1024
    x = (synth.randn(30, B, C, L) * 0.05).requires_grad_()
    w = (synth.randn(31, C, C, 5) * 0.02).requires_grad_()
    b = (synth.randn(32, C) * 0.01).requires_grad_()
    wj = (synth.randn(33, 1, C, 3) * 0.02).requires_grad_()
    bj = (synth.randn(34, 1) * 0.01).requires_grad_()
    y = F.leaky_relu(F.conv1d(x, w, b, padding=2), 0.2)
    j = F.conv1d(y, wj, bj, padding=1)
    r = synth.randn(35, *j.shape)
    (j * r).sum().backward()
    with torch.enable_grad():
        xg = x.detach().cuda().requires_grad_()
        wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
        wjg, bjg = wj.detach().cuda().requires_grad_(), bj.detach().cuda().requires_grad_()
        y32 = ag.DenseConvNCL.apply(xg, wg, bg, ag.WeightCache(), 2, True)
        jg = ag.MonoConv.apply(y32, wjg, bjg, 3, 1, False)
        assert rel_l2(jg, j) < 1e-4
        jg.backward(r.cuda())
    for name, got, ref in (("dx", xg.grad, x.grad), ("dw", wg.grad, w.grad), ("db", bg.grad, b.grad),
                           ("dwj", wjg.grad, wj.grad), ("dbj", bjg.grad, bj.grad)):
        e = rel_l2(got, ref)
        print("short L", L, name, e)
        assert e < 5e-3, (name, e)


@pytest.mark.parametrize("B,C,Co,L,s", [(2, 128, 128, 512, 4), (2, 128, 128, 128, 2), (3, 128, 128, 33, 4),
                                        (2, 64, 128, 100, 2)])
def test_strided_conv_block_backward(B, C, Co, L, s):
    """stride-s k7 conv + LeakyReLU as space-to-depth + tcgen05 conv (discriminator/multiscale.py:
    83-88): forward, input gradient (depth-to-space of the dgrad conv) and weight gradient (tap
    re-mapping of the wgrad GEMM) vs CPU autograd; the input carries one extra row like the
    filter bank's analysis output"""
    from music_synthesis_b200 import autograd as ag, grad_ops, ops
    x = synth.randn(40, B, C, L + 1).requires_grad_()
    w = (synth.randn(41, Co, C, 7) * 0.05).requires_grad_()
    b = (synth.randn(42, Co) * 0.1).requires_grad_()
    y = F.leaky_relu(F.conv1d(x[:, :, :L], w, b, stride=s, padding=3), 0.2)
    r = synth.randn(43, *y.shape)
    (y * r).sum().backward()
    xg = x.detach().cuda().requires_grad_()
    wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
    x32 = ag.PackBlk32.apply(xg)
    y32, _ = ag.StridedConvBlk.apply(x32, ops.pack_ncl(xg.detach()), wg, bg, ag.StridedCache(), s, L)
    yg = ag.UnpackBlk32.apply(y32)
    assert yg.shape == y.shape and rel_l2(yg, y) < 1e-3
    yg.backward(r.cuda())
    errs = {name: rel_l2(got, ref) for name, got, ref in
            (("dx", xg.grad, x.grad), ("dw", wg.grad, w.grad), ("db", bg.grad, b.grad))}
    print("strided", (B, C, Co, L, s), errs)
    # 1.4e-2 measured with bf16 AND fp16 backward operands alike (db, a pure fp32 sum, included):
    # ~3e-4 of the LeakyReLU masks differ because the forward ran on fp16 operands
    assert max(errs.values()) < 3e-2, errs


def test_filter_bank_ops_backward():
    """fixed Morlet bank: analysis (differentiable input) and synthesis (differentiable
    activations) vs CPU autograd on the same bank tensor"""
    from music_synthesis_b200 import autograd as ag, ops
    from music_synthesis_b200.audio.filterbank import FilterBank, linear_center_frequencies
    fb = FilterBank(22050, 128, linear_center_frequencies(0, 11025, 128), scaling_factors=0.05).to("cuda")
    bank = fb.filter_bank.cpu()
    B, L = 2, 300
    x = (synth.randn(44, B, 1, L) * 0.1).requires_grad_()
    a = F.conv1d(x, bank, padding=64)                      # (B,128,L+1)
    ra = synth.randn(45, *a.shape)
    (a * ra).sum().backward()
    xg = x.detach().cuda().requires_grad_()
    a32, _ = ag.BankAnalysis.apply(xg, fb)
    ag.UnpackBlk32.apply(a32).backward(ra.cuda())
    e = rel_l2(xg.grad, x.grad)
    print("bank analysis dx", e)
    assert e < 6e-3
    h = synth.randn(46, B, 128, L).requires_grad_()
    yb = F.conv_transpose1d(F.pad(h, (0, 1)), bank, padding=64)
    rb = synth.randn(47, *yb.shape)
    (yb * rb).sum().backward()
    hg = h.detach().cuda().requires_grad_()
    ys = ag.BankSynthesis.apply(ag.PackBlk32.apply(hg), ops.pack_ncl(hg.detach()), fb)
    assert rel_l2(ys, yb) < 1e-3
    ys.backward(rb.cuda())
    e = rel_l2(hg.grad, h.grad)
    print("bank synthesis dh", e)
    assert e < 6e-3
