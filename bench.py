#!/usr/bin/env python
"""bench.py -- headline benchmark of the stage-two hot path (BASELINE.json config 3):
MelGanGenerator inference, 256 clips x 128-bin mel x 256 frames -> 256 x 65536 samples
per GPU, clips sharded across ranks with no communication (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SAMPLE_RATE = 22050
FLOP_PER_SAMPLE = 409536          # 2 x MAC over conv / convT layers
CLIPS, MELS, FRAMES = 256, 128, 256
METRIC = "generated audio samples/sec"
UNIT = "samples/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"burst": p["bf16_tflops"], "sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.period_ms = os.environ.get("MSB_BENCH_SAMPLE_MS", "20")
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", self.period_ms],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port(clips, frames, steps, warmup):
    """Times the oracle port (the reference's algorithm, fp32 PyTorch on host cores)."""
    import torch
    from oracle import restate, synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = restate.melgan_generator_state(0)
    x = synth.mel_features(1, clips, frames)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            y = restate.melgan_generator(x, sd)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    samples = y.shape[0] * y.shape[-1]
    total = sum(times)
    return {"value": samples * len(times) / total, "unit": UNIT,
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle/restate.py melgan_generator, %d clips x %d frames, "
                      "%d timed passes, fp32, torch %s CPU" % (clips, frames, len(times),
                                                               torch.__version__),
            "ms_per_pass": 1e3 * total / len(times)}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path.  The reference
    is pure Python with no install recipe and un-installable dependencies, and
    /root/reference does not exist on the GPU box, so this times the oracle port (kind
    "port") with all host threads on a bounded sample of the same workload."""
    if rank != 0:
        return
    clips = 4
    t0 = time.perf_counter()
    base = cpu_port(clips, FRAMES, args.steps, max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_pass"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "x_realtime": base["value"] / SAMPLE_RATE,
        # the same config object as the B200 arm; what was actually timed (a bounded sample of
        # that workload: `clips` of its 256 clips per pass, throughput is per sample) is stated in
        # cpu_baseline.sample and below
        "config": workload_config(world),
        "reference_sample": "%d of the %d clips per step (CPU port, all host threads); samples/s "
                            "is size-independent: clips are processed independently" % (clips, CLIPS),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world):
    cfg = {
        "workload": "BASELINE config 3: MelGanGenerator inference, %d clips x %d-bin mel x "
                    "%d frames -> %d x %d samples per GPU, random-init weights "
                    "(N(0,0.02), zero bias), clips sharded by rank, no communication"
                    % (CLIPS, MELS, FRAMES, CLIPS, 256 * FRAMES),
        "clips_per_gpu": CLIPS, "global_clips": CLIPS * world, "mel_bins": MELS,
        "frames": FRAMES, "samples_per_clip": 256 * FRAMES,
        "l2": "activation working set per pass (>1 GB) exceeds the 126 MB L2; no explicit flush",
    }
    return cfg


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_over_ranks(torch, dist, world, dev, *vals):
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def bench_strong(torch, dist, gen, world, rank, dev, steps):
    """BASELINE config 3 as written: ONE global batch of 256 clips sharded by clip over the ranks
    (256/N per GPU, no communication).  Returns ms per step (max over ranks)."""
    import numpy as np
    from music_synthesis_b200.sharding import clip_shard
    lo, hi = clip_shard(CLIPS, rank, world)
    x = torch.from_numpy(np.random.RandomState(2000 + rank).standard_normal(
        (hi - lo, MELS, FRAMES)).astype(np.float32)).to(dev)
    with torch.no_grad():
        for _ in range(3):
            gen(x)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = _events(torch)
        e0.record()
        for _ in range(steps):
            gen(x)
        e1.record()
        torch.cuda.synchronize()
    (ms,) = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1) / steps)
    return {"global_clips": CLIPS, "clips_per_gpu": hi - lo, "ms_per_step": ms,
            "value": CLIPS * 256 * FRAMES / (ms * 1e-3), "unit": UNIT, "scaling": "strong"}


def bench_train(torch, dist, pair, global_batch, frames, world, rank, dev, cycles, warmup=3):
    """BASELINE configs 4 / 5: one training cycle = DiscriminatorTrainer.train + GeneratorTrainer.train
    (the reference's order, experiment/experiment.py:141-144) on a global batch sharded over the
    ranks; the gradient all-reduce (NCCL) of the network being stepped is inside the timed region.
    pair = "melgan" (cfg4: MelGanGenerator + MelGanDiscriminator, hinge) or "fb" (cfg5:
    FilterBankMultiScale pair, least squares, band dictionaries)."""
    from music_synthesis_b200 import _lib
    from music_synthesis_b200.experiment.init import weights_init
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import mel_gan_disc_loss, mel_gan_gen_loss
    per = max(1, global_batch // world)
    torch.manual_seed(10 + rank)            # Adam broadcasts rank 0's initial weights
    gsrc = torch.Generator(device=dev).manual_seed(100 + rank)
    feats = torch.randn(per, MELS, frames, device=dev, generator=gsrc) * 0.5 - 2.0
    real = torch.randn(per, 1, 256 * frames, device=dev, generator=gsrc) * 0.1
    kw = {}
    if pair == "fb":
        from music_synthesis_b200.generator.multiscale import FilterBankMultiScaleGenerator
        from music_synthesis_b200.discriminator.multiscale import FilterBankMultiScaleDiscriminator
        from music_synthesis_b200.audio.transform import fft_frequency_decompose
        from music_synthesis_b200.loss.loss import (least_squares_disc_loss,
                                                    least_squares_generator_loss)
        n = 256 * frames
        g = FilterBankMultiScaleGenerator(SAMPLE_RATE, MELS, frames, n, recompose=False).to(dev)
        d = FilterBankMultiScaleDiscriminator(n, SAMPLE_RATE, decompose=False,
                                              conditioning_channels=MELS).to(dev)
        with torch.no_grad():
            real = fft_frequency_decompose(real, n // 16)
        kw = {"d": least_squares_disc_loss, "g": least_squares_generator_loss}
    else:
        from music_synthesis_b200.generator.full import MelGanGenerator
        from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
        g = MelGanGenerator(frames, MELS).to(dev)
        d = MelGanDiscriminator().to(dev)
    g.apply(weights_init)
    d.apply(weights_init)
    g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, cuda_graph=True)
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, cuda_graph=True)
    if kw:
        d_tr.sub_loss, g_tr.sub_loss = kw["d"], kw["g"]
    with torch.enable_grad():
        for _ in range(warmup + 1):          # two eager calls per shape, then the capture
            d_tr.train(real, feats)
            g_tr.train(real, feats)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = _lib.launch_count()
        e0, e1 = _events(torch)
        e0.record()
        for _ in range(cycles):
            dl = d_tr.train(real, feats)["d_loss"]
            gl = g_tr.train(real, feats)["g_loss"]
        e1.record()
        torch.cuda.synchronize()
    # replayed graph nodes are not individual launch calls: add the captured counts
    launches = (_lib.launch_count() - l0) / cycles + d_tr.graph_launches + g_tr.graph_launches
    ms = e0.elapsed_time(e1) / cycles
    # the two all-reduces of a cycle (D gradients, G gradients) timed alone
    ar = 0.0
    if world > 1:
        dist.barrier()
        a0, a1 = _events(torch)
        a0.record()
        for _ in range(5):
            d_optim.all_reduce_grads()
            g_optim.all_reduce_grads()
        a1.record()
        torch.cuda.synchronize()
        ar = a0.elapsed_time(a1) / 5
    ms, ar = _max_over_ranks(torch, dist, world, dev, ms, ar)
    out = {"global_batch": per * world, "clips_per_gpu": per, "samples_per_clip": 256 * frames,
           "ms_per_cycle": ms, "clips_per_s": per * world / (ms * 1e-3), "allreduce_ms": ar,
           "allreduce_bytes": 4 * (g_optim.flat_grad.numel() + d_optim.flat_grad.numel()),
           "launches_per_cycle": launches, "cuda_graph": True, "d_loss": dl, "g_loss": gl,
           "scaling": "strong"}
    del d_tr, g_tr, g, d, g_optim, d_optim
    torch.cuda.empty_cache()
    return out


def bench_cfg1(torch, dev):
    """BASELINE config 1 on the GPU: batch 1 x 128 mel x 64 frames -> 16384 samples, latency."""
    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.experiment.init import weights_init
    torch.manual_seed(0)
    g = MelGanGenerator(64, MELS).eval()
    g.apply(weights_init)
    g = g.to(dev)
    x = torch.randn(1, MELS, 64, device=dev)
    with torch.no_grad():
        for _ in range(5):
            g(x)
        torch.cuda.synchronize()
        e0, e1 = _events(torch)
        e0.record()
        for _ in range(50):
            g(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    return {"gpu_ms": ms, "gpu_samples_per_s": 16384 / (ms * 1e-3)}


def bench_cfg2(torch, dev, hbm_gbs):
    """BASELINE config 2: Audio2Mel (STFT 1024 / hop 256 + 128-bin mel) on 64 x 16384 samples;
    HBM-bound: algorithmic bytes = 4*B*N in + 4*B*128*F out."""
    from music_synthesis_b200.feature.feature import Audio2Mel
    a2m = Audio2Mel(1024, 256, 1024, SAMPLE_RATE, 128).to(dev)
    out = {}
    for B in (64, 4096):
        a = torch.rand(B, 1, 16384, device=dev) * 2 - 1
        with torch.no_grad():
            for _ in range(5):
                m = a2m(a)
            torch.cuda.synchronize()
            e0, e1 = _events(torch)
            e0.record()
            for _ in range(50):
                m = a2m(a)
            e1.record()
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        nbytes = 4 * B * 16384 + 4 * m.numel()
        out["b%d" % B] = {"us": us, "samples_per_s": B * 16384 / (us * 1e-6),
                          "achieved_gbs": nbytes / (us * 1e-6) / 1e9,
                          "hbm_frac": nbytes / (us * 1e-6) / 1e9 / hbm_gbs}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true",
                    help="skip the strong-scaling / training / cfg1 / cfg2 side measurements")
    ap.add_argument("--e2e-chunk", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--e2e-edge", type=int, default=-1, help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from music_synthesis_b200 import _lib
    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.experiment.init import weights_init

    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    clips = args.clips
    # ---- model + synthetic inputs (each rank: its own contiguous shard of clips)
    # random-init weights by the reference's own contract (experiment/init.py: N(0, 0.02) weights,
    # zero biases); nothing on this arm touches oracle/ (only cpu_baseline() below does)
    torch.manual_seed(0)
    gen = MelGanGenerator(FRAMES, MELS).eval()
    gen.apply(weights_init)
    gen = gen.to(dev)
    # weak scaling: the global synthetic batch has clips*world clips; this rank's contiguous
    # shard (no data-path collective) is regenerated locally from its own seed
    from music_synthesis_b200.sharding import clip_shard
    lo, hi = clip_shard(clips * world, rank, world)
    assert hi - lo == clips
    import numpy as np
    x_host = torch.from_numpy(np.random.RandomState(1000 + rank).standard_normal(
        (clips, MELS, FRAMES)).astype(np.float32)).pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    y_host = torch.empty((clips, 1, 256 * FRAMES), dtype=torch.float32).pin_memory()
    samples_per_step = clips * 256 * FRAMES

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------
    with torch.no_grad():
        # clocks are sampled from before the warm-up (nvidia-smi needs ~0.5 s to deliver its first
        # row) through the timed region; a short timed region is followed by an UNTIMED
        # continuation of the same step until enough rows exist -- all of it under this load
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            y = gen(x_dev)
        sync_all()
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = gen(x_dev)
        e1.record()
        sync_all()
        launches = _lib.launch_count() - launches0
        ms = e0.elapsed_time(e1)
        if rank == 0:
            t_hold = time.time()
            while len(sampler.rows) < 10 and time.time() - t_hold < 4.0:
                y = gen(x_dev)
                torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end through the module with HOST buffers ---------------------
        # the public host-to-host call: pinned features in, pinned waveform out, every step
        # copies its 33.5 MB of inputs H2D and its 67 MB of audio D2H inside the timed region
        ekw = {"chunk_clips": args.e2e_chunk} if args.e2e_chunk > 0 else {}
        if args.e2e_edge >= 0:
            ekw["edge_clips"] = args.e2e_edge
        for _ in range(2):
            gen.generate(x_host, out=y_host, **ekw)
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            gen.generate(x_host, out=y_host, **ekw)
        f1.record()
        sync_all()
        ms_e2e = f0.elapsed_time(f1)

        # ---- dominant kernel, timed alone on this stream: the fused ResidualStack at C=128
        # (stage 2: largest single kernel of the step).  Same shapes as inside the step.
        from music_synthesis_b200 import ops as _ops
        stack = gen.main[8]
        blob = _ops.resstack_pack_weights(list(stack.parameters()), 128)
        x32 = torch.randn((clips, 16, 64 * FRAMES, 8), device=dev) * 0.1
        for _ in range(2):
            _ops.resstack_fwd(x32, blob, [1, 3, 9], want16=True, want32=False)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        k0.record()
        for _ in range(reps):
            _ops.resstack_fwd(x32, blob, [1, 3, 9], want16=True, want32=False)
        k1.record()
        torch.cuda.synchronize()
        us_stack = k0.elapsed_time(k1) * 1e3 / reps
        del x32

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- what else the north star asks for, on every N (all ranks take part):
    extras = {}
    if not args.headline_only:
        extras["strong"] = bench_strong(torch, dist, gen, world, rank, dev, max(5, args.steps // 2))
        del gen
        torch.cuda.empty_cache()
        extras["train_cfg4"] = bench_train(torch, dist, "melgan", 32, 32, world, rank, dev, 10)
        extras["train_cfg5"] = bench_train(torch, dist, "fb", 64, 256, world, rank, dev, 5)
        if rank == 0:
            extras["cfg1"] = bench_cfg1(torch, dev)
            extras["cfg2"] = bench_cfg2(torch, dev, peaks()["hbm_gbs"])

    if rank == 0:
        pk = peaks()
        total_samples = samples_per_step * world * args.steps
        value = total_samples / (ms * 1e-3)
        e2e_value = total_samples / (ms_e2e * 1e-3)
        step_ms = ms / args.steps
        tflops_per_gpu = FLOP_PER_SAMPLE * samples_per_step / (step_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate (tcgen05 kind::f16), f32 residual stream",
            "data": "synthetic",
            "x_realtime": value / SAMPLE_RATE,
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "x_realtime": e2e_value / SAMPLE_RATE,
                    "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": y_host.numel() * 4},
            "gpu_launches": launches,
            # dominant kernel (largest single launch of the step), timed alone with CUDA
            # events above; algorithmic FLOPs = 6 convs x 2*3*128^2 per output row.
            "roofline": {
                "bound": "tensor", "kernel": "resstack_pair_kernel<128> (fused ResidualStack, stage 2, CTA pairs)",
                "achieved": 6 * 2 * 3 * 128 * 128 * clips * 64 * FRAMES / (us_stack * 1e-6) / 1e12,
                "peak": pk["burst"], "unit": "TFLOP/s",
                "frac": 6 * 2 * 3 * 128 * 128 * clips * 64 * FRAMES / (us_stack * 1e-6) / 1e12 / pk["burst"],
                "peak_source": pk["source"] + ", burst figure (kernel timed alone)",
                "us_per_launch": us_stack,
                # dram__bytes_read+write of this kernel from profiles/r02_ncu_stacks_summary.tsv
                # (2.155 + 1.045 GB per 256-clip launch of the CTA-pair kernel, scaled to this
                # launch's clip count); the algorithmic bytes are 4C in + 2C out per row = 3.22 GB
                "traffic": 3.2007e9 * clips / 256,
                "traffic_source": "ncu constant (dram__bytes_read+write of this kernel at 256 clips, "
                                  "profiles/r02_ncu_stacks_summary.tsv), scaled by clip count; not "
                                  "measured in this run",
            },
            # the whole step (15 kernels) against the same roofline: algorithmic generator FLOPs
            # (409 536 per sample) / CUDA-event step time, sustained peak (seconds-long step)
            "roofline_step": {
                "bound": "tensor", "achieved": tflops_per_gpu, "peak": pk["sustained"],
                "unit": "TFLOP/s", "frac": tflops_per_gpu / pk["sustained"],
                "frac_of_burst": tflops_per_gpu / pk["burst"], "peak_source": pk["source"],
                "flop_per_sample": FLOP_PER_SAMPLE,
            },
            "clocks": clocks,
        }
        if "strong" in extras:
            # x of one GPU: this run's own 256-clips-on-one-GPU time is the weak step above
            extras["strong"]["x_of_one_gpu"] = step_ms / extras["strong"]["ms_per_step"]
        line.update(extras)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port(8, FRAMES, 2, 1)
            line["cpu_baseline"].pop("ms_per_pass", None)
            if "cfg1" in line:
                c1 = cpu_port(1, 64, 5, 2)
                line["cfg1"].update({"cpu_port_ms": c1["ms_per_pass"], "cpu_cores": c1["cores"],
                                     "cpu_samples_per_s": c1["value"]})
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
