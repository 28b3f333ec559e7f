"""-m gpu: the fused upsampling stage (ConvTranspose1d(2C->C, 4, 2, 1) + LeakyReLU + ResidualStack
[+ k7 tail + tanh], csrc/upstack.cu) against the same layers in PyTorch on the host with the
kernel's rounding points (16-bit operands, wide accumulate, fp32 residual stream), at tile-edge
lengths, many tiles per CTA, and against the unfused device schedule (generator/full.py:35-44)."""
import pytest
import torch
from torch.nn import functional as F

from tests.gpu_util import rel_l2, rnd16, randn
from tests.test_gpu_kernels import _stack_emulated

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from music_synthesis_b200 import ops as _ops
    return _ops


def _stage_params(seed, C):
    from oracle import synth
    sd = synth.residual_stack_state(seed, C)
    wt = randn(seed + 1, 2 * C, C, 4, scale=0.05)
    bt = randn(seed + 2, C, scale=0.1)
    params = [wt.cuda(), bt.cuda()]
    for a in range(3):
        for c in range(2):
            params += [sd[f"s.main.{a}.main.{c}.weight"].cuda(), sd[f"s.main.{a}.main.{c}.bias"].cuda()]
    return sd, wt, bt, params


def _stage_emulated(x, sd, wt, bt):
    up = F.leaky_relu(F.conv_transpose1d(rnd16(x).double(), rnd16(wt).double(), bt.double(),
                                         stride=2, padding=1), 0.2)
    return _stack_emulated(up.float(), sd)


# lin chosen around the tile geometry: C=64 stores V=476 rows per tile, C=32 V=988 (982 with tail)
@pytest.mark.parametrize("C,B,lin", [(64, 1, 100), (64, 2, 238), (64, 2, 239), (64, 1, 1203),
                                     (32, 1, 64), (32, 2, 494), (32, 3, 495), (32, 1, 2600)])
def test_fused_upsampling_stage(ops, C, B, lin):
    sd, wt, bt, params = _stage_params(300 + C, C)
    x = randn(310 + C, B, 2 * C, lin, scale=0.5)
    ref = _stage_emulated(x, sd, wt, bt).float()
    blob = ops.upstack_pack_weights(params, C)
    y16, y32 = ops.upstack_fwd(ops.pack_ncl(x.cuda()), blob, [1, 3, 9], want16=True, want32=True)
    torch.cuda.synchronize()
    got = ops.unpack_blk32(y32).cpu()
    assert got.shape == (B, C, 2 * lin)
    per_row = (got - ref).abs().amax(dim=(0, 1))
    assert rel_l2(got, ref) < 2e-5, (rel_l2(got, ref), torch.nonzero(per_row > 1e-3).flatten()[:20])
    assert torch.equal(ops.unpack_blk16(y16).cpu(), rnd16(got))


@pytest.mark.parametrize("B,lin", [(1, 40), (2, 491), (2, 492), (5, 3000)])
def test_fused_last_stage_with_tail(ops, B, lin):
    C = 32
    sd, wt, bt, params = _stage_params(400, C)
    tw, tb = randn(403, 1, 32, 7, scale=0.2), randn(404, 1, scale=0.1)
    x = randn(410, B, 2 * C, lin, scale=0.5)
    stack = _stage_emulated(x, sd, wt, bt)
    ref = torch.tanh(F.conv1d(stack, tw.double(), tb.double(), padding=3)).float()
    blob = ops.upstack_pack_weights(params, C)
    got = ops.upstack_fwd(ops.pack_ncl(x.cuda()), blob, [1, 3, 9], tail=(tw.cuda(), tb.cuda())).cpu()
    assert got.shape == (B, 1, 2 * lin)
    assert rel_l2(got, ref) < 2e-5


def test_fused_stage_many_tiles_matches_unfused_device_schedule(ops):
    """More tiles than CTAs (buffer alternation, ring wrap, barrier parities beyond the first
    iterations): the fused launch against ConvTranspose (conv kernel) -> fused stack."""
    for C, B, lin in ((64, 12, 4096), (32, 6, 16384)):
        sd, wt, bt, params = _stage_params(500 + C, C)
        x16 = ops.pack_ncl(randn(510 + C, B, 2 * C, lin, scale=0.5).cuda())
        y16, y32 = ops.upstack_fwd(x16, ops.upstack_pack_weights(params, C), [1, 3, 9],
                                   want16=True, want32=True)
        d = ops.conv_desc(ops.MS_CONVT, B, 2 * C, C, lin, 4, 1, 1, 2, leaky=True)
        _, u32 = ops.conv_fwd(d, x16, ops.pack_conv_weight(d, params[0]), params[1],
                              want16=False, want32=True)
        r16, r32 = ops.resstack_fwd(u32, ops.resstack_pack_weights(params[2:], C), [1, 3, 9],
                                    want16=True, want32=True)
        torch.cuda.synchronize()
        per_clip = ((y32 - r32).double().flatten(1).norm(dim=1) /
                    r32.double().flatten(1).norm(dim=1))
        assert per_clip.max().item() < 1e-5, per_clip.tolist()


def test_fused_stage_rejects_even_dilations(ops):
    from music_synthesis_b200._lib import MsbError
    sd, wt, bt, params = _stage_params(600, 32)
    blob = ops.upstack_pack_weights(params, 32)
    x16 = ops.pack_ncl(randn(601, 1, 64, 64).cuda())
    with pytest.raises(MsbError):
        ops.upstack_fwd(x16, blob, [1, 2, 9])


@pytest.mark.parametrize("C,B,lin", [(64, 3, 3), (64, 1, 9), (32, 2, 5), (32, 5, 17)])
def test_fused_stage_on_clips_shorter_than_the_halo(ops, C, B, lin):
    """both clip edges inside one tile, fewer input rows than the stack's receptive field"""
    sd, wt, bt, params = _stage_params(700 + C, C)
    x = randn(710 + lin, B, 2 * C, lin, scale=0.5)
    ref = _stage_emulated(x, sd, wt, bt).float()
    _, y32 = ops.upstack_fwd(ops.pack_ncl(x.cuda()), ops.upstack_pack_weights(params, C), [1, 3, 9],
                             want16=False, want32=True)
    assert rel_l2(ops.unpack_blk32(y32).cpu(), ref) < 2e-5


@pytest.mark.parametrize("C", [64, 32])
def test_fused_stage_bf16_operands(ops, C):
    """the bf16 operand mode of the fused stage against the same layers with bf16-rounded
    operands (wide accumulate, fp32 stream)"""
    from music_synthesis_b200._lib import MS_BF16
    B, lin = 2, 300
    sd, wt, bt, params = _stage_params(800 + C, C)
    x = randn(810 + C, B, 2 * C, lin, scale=0.5)
    up = F.leaky_relu(F.conv_transpose1d(rnd16(x, "bf16").double(), rnd16(wt, "bf16").double(),
                                         bt.double(), stride=2, padding=1), 0.2)
    ref = _stack_emulated(up.float(), sd, operand="bf16").float()
    blob = ops.upstack_pack_weights(params, C, operand=MS_BF16)
    _, y32 = ops.upstack_fwd(ops.pack_ncl(x.cuda(), operand=MS_BF16), blob, [1, 3, 9],
                             operand=MS_BF16, want16=False, want32=True)
    assert rel_l2(ops.unpack_blk32(y32).cpu(), ref) < 2e-5
