#!/usr/bin/env python
"""Data-parallel training check, run under torchrun on N GPUs (NCCL): every rank trains on its
shard of a global batch; rank 0 also trains a second, non-distributed copy on the WHOLE batch.
After a D step and a G step the two must hold the same weights (gradient all-reduce = sum over
ranks x 1/N of batch-mean losses = the global batch mean).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    print("rank", rank, "init ok", flush=True)
    import music_synthesis_b200  # noqa: F401
    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
    from music_synthesis_b200.experiment.init import weights_init
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import (mel_gan_disc_loss, mel_gan_gen_loss,
                                                least_squares_disc_loss,
                                                least_squares_generator_loss)
    T, per = 8, 2

    def build(distributed, graph):
        # data-parallel replicas draw DIFFERENT initial weights per rank on purpose: Adam's
        # constructor broadcasts rank 0's; the single-process copy re-draws rank 0's
        torch.manual_seed(1000 * rank if distributed else 0)
        g = MelGanGenerator(T, 128).cuda()
        d = MelGanDiscriminator().cuda()
        g.apply(weights_init)
        d.apply(weights_init)
        go = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9), distributed=distributed)
        do = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9), distributed=distributed)
        dt = DiscriminatorTrainer(g, go, d, do, mel_gan_disc_loss, least_squares_disc_loss, cuda_graph=graph)
        gt = GeneratorTrainer(g, go, d, do, mel_gan_gen_loss, least_squares_generator_loss, cuda_graph=graph)
        return g, d, dt, gt

    gen = torch.Generator(device="cuda").manual_seed(7)
    feats = torch.randn(world * per, 128, T, device="cuda", generator=gen) * 0.5 - 2.0
    real = torch.randn(world * per, 1, 256 * T, device="cuda", generator=gen) * 0.1
    lo = rank * per
    for graph in (False, True):
        g, d, dt, gt = build(True, graph)
        steps = 4 if graph else 1
        for i in range(steps):
            rd = dt.train(real[lo:lo + per], feats[lo:lo + per])
            rg = gt.train(real[lo:lo + per], feats[lo:lo + per])
            print('rank', rank, 'graph', graph, 'step', i, 'ok', flush=True)
        if rank == 0:
            g2, d2, dt2, gt2 = build(False, False)
            for _ in range(steps):
                rd2 = dt2.train(real, feats)
                rg2 = gt2.train(real, feats)
            worst = 0.0
            for (k, a), (_, b) in list(zip(g.state_dict().items(), g2.state_dict().items())) + \
                    list(zip(d.state_dict().items(), d2.state_dict().items())):
                # first Adam steps move each weight by ~lr: compare the update, not the weight
                worst = max(worst, float((a - b).abs().max()))
            print("graph=%s steps=%d  d_loss %.6f (dp, local shard) vs %.6f (whole batch)  "
                  "max |w_dp - w_single| = %.3e (lr = 1e-4)" % (graph, steps, rd["d_loss"], rd2["d_loss"], worst))
            frac = 0.0
            n = 0
            for (k, a), (_, b) in zip(g.state_dict().items(), g2.state_dict().items()):
                frac += float(((a - b).abs() > 0.5e-4).sum())
                n += a.numel()
            print("   generator weights differing by more than lr/2: %.4f %%" % (100.0 * frac / n))
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
