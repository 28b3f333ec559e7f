#!/usr/bin/env python
"""BASELINE.json config 4: full GAN training step (MelGanGenerator + 3-scale MelGanDiscriminator,
feature-matching + hinge losses, Adam 1e-4 (0.5, 0.9)), global batch 32 clips x 8192 samples
(T = 32 frames), data-parallel over the ranks (gradient all-reduce = one NCCL call per step on
the flat gradient buffer).  One "cycle" = DiscriminatorTrainer.train + GeneratorTrainer.train,
the reference's order (experiment/experiment.py:141-144).

    python tools/train_bench.py [--steps K] [--warmup W] [--batch 32] [--frames 32]
    python -m torch.distributed.run --nproc-per-node N ... tools/train_bench.py --gpus N

Prints one JSON line (rank 0): cycles/s, clips/s, ms per cycle (CUDA events, max over ranks),
kernel launches per cycle, host time per cycle.  Strong scaling: the global batch is fixed.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="GLOBAL batch (clips)")
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--graph", type=int, default=1, help="1: CUDA-graph the steps (default)")
    ap.add_argument("--pair", default="melgan", choices=["melgan", "fb"],
                    help="melgan: cfg4 (MelGAN pair, hinge); fb: cfg5 (filter-bank multiscale pair, "
                         "least squares, 65536-sample clips: use --batch 64 --frames 256)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import music_synthesis_b200  # noqa: F401
    from music_synthesis_b200 import _lib
    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
    from music_synthesis_b200.experiment.init import weights_init
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import mel_gan_disc_loss, mel_gan_gen_loss

    torch.manual_seed(0)                       # same init on every rank
    per = args.batch // world
    gen = torch.Generator(device="cuda").manual_seed(1 + rank)
    feats = torch.randn(per, 128, args.frames, device="cuda", generator=gen) * 0.5 - 2.0
    real = torch.randn(per, 1, 256 * args.frames, device="cuda", generator=gen) * 0.1
    sub = {}
    if args.pair == "fb":
        # experiment/multiscale.py:16-67: band dictionaries in, least-squares sub-losses
        from music_synthesis_b200.generator.multiscale import FilterBankMultiScaleGenerator
        from music_synthesis_b200.discriminator.multiscale import FilterBankMultiScaleDiscriminator
        from music_synthesis_b200.audio.transform import fft_frequency_decompose
        from music_synthesis_b200.loss.loss import (least_squares_disc_loss,
                                                    least_squares_generator_loss)
        n = 256 * args.frames
        g = FilterBankMultiScaleGenerator(22050, 128, args.frames, n, recompose=False).cuda()
        d = FilterBankMultiScaleDiscriminator(n, 22050, decompose=False,
                                              conditioning_channels=128).cuda()
        with torch.no_grad():
            real = fft_frequency_decompose(real, n // 16)
        sub = {"d": least_squares_disc_loss, "g": least_squares_generator_loss}
    else:
        g = MelGanGenerator(args.frames, 128).cuda()
        d = MelGanDiscriminator().cuda()
    g.apply(weights_init)
    d.apply(weights_init)
    g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, cuda_graph=bool(args.graph))
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, cuda_graph=bool(args.graph))
    if sub:
        d_tr.sub_loss, g_tr.sub_loss = sub["d"], sub["g"]

    def cycle():
        a = d_tr.train(real, feats)
        b = g_tr.train(real, feats)
        return a["d_loss"], b["g_loss"]

    for _ in range(args.warmup):
        losses = cycle()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.lib().ms_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        losses = cycle()
    e1.record()
    torch.cuda.synchronize()
    host = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    launches = (_lib.lib().ms_launch_count() - l0) / args.steps
    if rank == 0:
        per_cycle = ms / args.steps
        print(json.dumps({
            "metric": "GAN training cycles/sec (D step + G step)", "value": 1000.0 / per_cycle,
            "unit": "cycles/s", "clips_per_s": args.batch * 1000.0 / per_cycle, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_cycle,
            "host_ms_per_step": host * 1000.0 / args.steps, "gpu_launches_per_step": launches,
            "higher_is_better": True, "scaling": "strong", "dtype": "fp16 fwd / bf16 bwd operands, fp32 accumulate",
            "data": "synthetic", "cuda_graph": bool(args.graph), "d_loss": losses[0], "g_loss": losses[1],
            "config": {"workload": "%s train cycle, global batch %d x %d samples, "
                                   "Adam(1e-4,(0.5,0.9)), DP x%d"
                                   % ("cfg5: FilterBankMultiScaleGenerator + FilterBankMultiScale"
                                      "Discriminator (least squares)" if args.pair == "fb" else
                                      "cfg4: MelGanGenerator + MelGanDiscriminator (hinge)",
                                      args.batch, 256 * args.frames, world)}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
