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
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
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
        "config": workload_config(world, sample_clips=clips),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, sample_clips=None):
    cfg = {
        "workload": "BASELINE config 3: MelGanGenerator inference, %d clips x %d-bin mel x "
                    "%d frames -> %d x %d samples per GPU, random-init weights "
                    "(N(0,0.02), zero bias), clips sharded by rank, no communication"
                    % (CLIPS, MELS, FRAMES, CLIPS, 256 * FRAMES),
        "clips_per_gpu": CLIPS, "global_clips": CLIPS * world, "mel_bins": MELS,
        "frames": FRAMES, "samples_per_clip": 256 * FRAMES,
        "l2": "activation working set per pass (>1 GB) exceeds the 126 MB L2; no explicit flush",
    }
    if sample_clips is not None:
        cfg["cpu_sample_clips"] = sample_clips
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=0, help=argparse.SUPPRESS)
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
                "bound": "tensor", "kernel": "resstack_kernel<128> (fused ResidualStack, stage 2)",
                "achieved": 6 * 2 * 3 * 128 * 128 * clips * 64 * FRAMES / (us_stack * 1e-6) / 1e12,
                "peak": pk["burst"], "unit": "TFLOP/s",
                "frac": 6 * 2 * 3 * 128 * 128 * clips * 64 * FRAMES / (us_stack * 1e-6) / 1e12 / pk["burst"],
                "peak_source": pk["source"] + ", burst figure (kernel timed alone)",
                "us_per_launch": us_stack,
                # dram__bytes_read+write of this kernel from profiles/r01_ncu_final_summary.tsv
                # (776.7 MB per 64-clip launch, scaled to this launch's clip count); the
                # algorithmic bytes are 4C in + 2C out per row = 805 MB per 64 clips
                "traffic": 776.7e6 * clips / 64,
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
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port(8, FRAMES, 2, 1)
            line["cpu_baseline"].pop("ms_per_pass", None)
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
