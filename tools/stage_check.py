"""Debug: per-stage comparison of the generator schedule (re-built from ops) vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from music_synthesis_b200 import ops
from oracle import restate, synth

torch.set_grad_enabled(False)
B, T = int(sys.argv[1]), int(sys.argv[2])
sd = restate.randomize_biases(restate.melgan_generator_state(3), 1003)
x = synth.mel_features(5, B, T)

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

lre = lambda t: F.leaky_relu(t, 0.2)
# oracle stage outputs
r = lre(F.conv1d(F.pad(x, (3, 3), mode="reflect"), sd["main.1.weight"], sd["main.1.bias"]))
ref = {"first": r}
for i, (ct, st, s, p) in enumerate(restate.MELGAN_UPSAMPLERS):
    r = lre(F.conv_transpose1d(r, sd[f"main.{ct}.weight"], sd[f"main.{ct}.bias"], stride=s, padding=p))
    ref[f"up{i}"] = r
    r = restate.residual_stack(r, sd, f"main.{st}")
    ref[f"stack{i}"] = r

cu = {k: v.cuda() for k, v in sd.items()}
x16 = ops.pack_ncl(x.cuda(), 3, 1)
d = ops.conv_desc(ops.MS_CONV, B, 128, 512, T + 6, 7, 1, 0, leaky=True)
h16, h32 = ops.conv_fwd(d, x16, ops.pack_conv_weight(d, cu["main.1.weight"]), cu["main.1.bias"], want16=True, want32=True)
print("first", rel(ops.unpack_blk32(h32), ref["first"]))
L = T
chans = {3: (512, 256, 16), 6: (256, 128, 16), 9: (128, 64, 4), 12: (64, 32, 4)}
for i, (ct, st, s, p) in enumerate(restate.MELGAN_UPSAMPLERS):
    cin, cout, k = chans[ct]
    d = ops.conv_desc(ops.MS_CONVT, B, cin, cout, L, k, 1, p, s, leaky=True)
    u16, u32 = ops.conv_fwd(d, h16, ops.pack_conv_weight(d, cu[f"main.{ct}.weight"]), cu[f"main.{ct}.bias"], want16=True, want32=True)
    L *= s
    got = ops.unpack_blk32(u32)
    e = rel(got, ref[f"up{i}"])
    print(f"up{i} (L={L})", e)
    if e > 1e-2:
        dd = (got.cpu() - ref[f"up{i}"]).abs().amax(dim=(0, 1))
        bad = torch.nonzero(dd > 1e-3 * ref[f"up{i}"].abs().max()).flatten()
        print("   bad rows:", bad[:10].tolist(), "...", bad[-10:].tolist(), "count", bad.numel())
    # stack on the ORACLE's upsampler output (isolates the stack)
    if ops.resstack_supported(cout):
        params = []
        for a in range(3):
            for c in range(2):
                params += [cu[f"main.{st}.main.{a}.main.{c}.weight"], cu[f"main.{st}.main.{a}.main.{c}.bias"]]
        blob = ops.resstack_pack_weights(params, cout)
        rin = ref[f"up{i}"]
        x32 = rin.view(B, cout // 8, 8, L).permute(0, 1, 3, 2).contiguous().cuda()
        y16, y32 = ops.resstack_fwd(x32, blob, [1, 3, 9], want16=True, want32=True)
        got = ops.unpack_blk32(y32)
        e = rel(got, ref[f"stack{i}"])
        print(f"stack{i} C={cout}", e)
        if e > 1e-2:
            dd = (got.cpu() - ref[f"stack{i}"]).abs().amax(dim=(0, 1))
            bad = torch.nonzero(dd > 1e-3 * ref[f"stack{i}"].abs().max()).flatten()
            print("   bad rows:", bad[:10].tolist(), "...", bad[-10:].tolist(), "count", bad.numel())
        h16 = ops.pack_ncl(ref[f"stack{i}"].cuda())
    else:
        h16 = ops.pack_ncl(ref[f"stack{i}"].cuda())
