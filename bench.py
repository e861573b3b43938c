#!/usr/bin/env python
"""Benchmark of the gym_dockauv step hot path on B200 (BASELINE.json metric: env-steps/s, ObstaclesDocking3d,
BlueROV2, 64-ray radar, 5 capsules + 3 spheres -- config C4: 1,048,576 envs per GPU, weak scaling to C5).

    python bench.py [--gpus N] [--steps K] [--warmup W]             # our arm (CUDA, through the C ABI)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU arm: the oracle port on host cores

One "step" = one batched env.step() over every env of the rank (default layout: four launches -- dynamics, cull, rays,
finish -- for each half of the batch, the halves on two streams).  Rank 0 prints ONE JSON line.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  The per-step
working set (~0.9 GB at 1M envs) is far larger than L2 (126 MB), so no explicit L2 flush is needed (stated in
config.l2).  Actions are synthetic i.i.d. U(-1,1) float32, pre-generated on the device for `value`; the `e2e`
number goes through env.step_host() with pinned HOST actions in and HOST obs/reward/done out every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_ENV_STEP = 898      # algorithmic HBM bytes per env-step for C4 (FP64 SoA, f32 obs/actions)
FLOPS_PER_ENV_STEP = 17700    # algorithmic flops per env-step for C4
HBM_FALLBACK_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback
# dram__bytes_read.sum + dram__bytes_write.sum of the eight launches (two halves x four) of ONE step over 1,048,576 envs
# (FP64) from the committed `ncu --set full` capture profiles/r01/v13_pipeline_ncu_raw.csv (per launch in that file)
NCU_TRAFFIC_BYTES_1M_F64 = 1756.4e6
LAUNCH_NAMES = ("dynamics", "cull", "rays", "finish")
# sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active per launch, same capture (fraction of the FP64 pipe's
# issue slots actually used -- the executed counterpart of the algorithmic `pipe.frac`)
NCU_FP64_PIPE_BUSY = {"dynamics": 0.496, "cull": 0.031, "rays": 0.367, "finish": 0.105}
WORKLOAD = ("C4: ObstaclesDocking3d, BlueROV2, 64-ray radar, 5 capsules + 3 spheres, random actions U(-1,1) f32, "
            "auto-reset of finished envs")
SCENARIO = "ObstaclesDocking3d"
N_SYNTH_SPHERES = 3


def workload_config():
    from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
    cfg = dict(BASE_CONFIG)
    cfg["radar"] = dict(RADAR_64)
    return cfg


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(n_envs, seconds, n_threads=0, steps_cap=10 ** 9, warmup=1):
    """Times the oracle port (oracle/dockauv_oracle.c, OpenMP over envs) on this host's cores."""
    from oracle import oracle as orc
    cfg = workload_config()
    bo = orc.BatchOracle(cfg, SCENARIO, n_envs, seed=0, n_extra_spheres=N_SYNTH_SPHERES, n_threads=n_threads)
    rng = np.random.default_rng(1)
    pool = [rng.uniform(-1, 1, (n_envs, 6)).astype(np.float32) for _ in range(4)]
    for i in range(warmup):
        bo.step(pool[i % 4])
    t0 = time.perf_counter()
    k = 0
    while k < steps_cap and (time.perf_counter() - t0 < seconds or k < 2):
        bo.step(pool[k % 4])
        k += 1
    dt = time.perf_counter() - t0
    threads = n_threads if n_threads > 0 else orc.lib().orc_max_threads()
    return n_envs * k / dt, threads, k, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as orc
    n = args.ref_envs
    cfg = workload_config()
    bo = orc.BatchOracle(cfg, SCENARIO, n, seed=0, n_extra_spheres=N_SYNTH_SPHERES)
    rng = np.random.default_rng(1)
    pool = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(4)]
    for i in range(args.warmup):
        bo.step(pool[i % 4])
    t0 = time.perf_counter()
    for k in range(args.steps):
        bo.step(pool[k % 4])
    dt = time.perf_counter() - t0
    threads = orc.lib().orc_max_threads()
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "envs_per_step": n, "sample": f"{n} envs per step (bounded sample of the 1,048,576-env workload)"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n} envs x {args.steps} steps, oracle/dockauv_oracle.c with OpenMP on {threads} threads"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from gym_dockauv_b200 import envs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL writes its version banner (and any debug output) to STDOUT by default; stdout carries only the one
        # JSON line, so NCCL's log goes to stderr
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, the image's default)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_config()
    N = args.envs_per_gpu
    env = envs.ObstaclesDocking3d(cfg, num_envs=N, device=dev, precision=args.precision, seed=args.seed,
                                  env_id0=rank * N, n_synthetic_spheres=N_SYNTH_SPHERES, layout=args.layout)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = [torch.rand(N, env.n_actions, device=dev, generator=gen) * 2 - 1 for _ in range(args.action_pool)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # burn-in: brings the batch from "every env just reset" to a mixed episode-age distribution (episodes last
    # ~100 steps under random actions), so the timed steps see the steady-state mix of ray hits and resets
    for k in range(args.burn_in):
        env.step(pool[k % len(pool)])
    for k in range(args.warmup):
        env.step(pool[k % len(pool)])
    stats_t = env.stats_tensor()
    env.clear_stats()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = env.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    side = torch.cuda.Stream(dev)
    n_reduces = 0
    evs[0].record()
    for k in range(args.steps):
        env.step(pool[k % len(pool)])
        evs[k + 1].record()
        if world > 1 and (k + 1) % args.rollout == 0:
            # per-rollout episode statistics: one small NCCL all-reduce on a side stream
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                red = env.stats_tensor().clone()      # folds the per-CTA replicas on the side stream first
                dist.all_reduce(red)
            n_reduces += 1
    torch.cuda.current_stream(dev).wait_stream(side)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)])
    launches = env.launch_count() - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    stats = env.get_stats()
    # per-launch durations (CUDA events recorded by the library between the launches of a step, on the launching
    # stream), in a separate short pass so that the marks do not sit inside the timed region above
    env.enable_timing(True)
    per_launch = []
    for k in range(min(32, args.steps)):
        env.step(pool[k % len(pool)])
        per_launch.append(env.last_step_ms()[1])
    env.enable_timing(False)
    launch_ms = np.array(per_launch).mean(axis=0) if per_launch and per_launch[0] else np.zeros(0)

    # ---- end to end through the public API with host buffers (rank-local, then max over ranks)
    e2e_steps = max(3, min(args.e2e_steps, args.steps))
    host_pool = [torch.empty(N, env.n_actions, dtype=torch.float32).pin_memory() for _ in range(2)]
    for hp, dp in zip(host_pool, pool):
        hp.copy_(dp)
    host_np = [hp.numpy() for hp in host_pool]
    env.step_host(host_np[0])
    env.step_host(host_np[1])
    barrier()
    t0 = time.perf_counter()
    sink = 0.0
    for k in range(e2e_steps):
        o, r, d, _ = env.step_host(host_np[k % 2])
        sink += float(r[0]) + float(o[0, 0]) + float(d[0])      # results really are in host memory
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s_max = float(t.item())
    esz = 8 if args.precision == "f64" else 4
    h2d = N * env.n_actions * 4
    d2h = N * (env.n_observations * 4 + esz + 1 + 1)

    if rank == 0:
        value = world * N * args.steps / (total_ms_max * 1e-3)
        kern_ms = float(step_ms.mean())
        per_gpu_rate = N / (kern_ms * 1e-3)
        hbm_peak, hbm_src = measured_hbm_peak()
        achieved_gbs = per_gpu_rate * BYTES_PER_ENV_STEP / 1e9
        fp64_peak = fp32_peak = None
        try:
            import ctypes as C
            from gym_dockauv_b200 import _capi
            a, b = C.c_double(), C.c_double()
            _capi.check(_capi.load().dockauv_measure_peaks(local_rank, C.byref(a), C.byref(b), None))
            fp64_peak, fp32_peak = a.value, b.value
        except Exception as ex:  # noqa: BLE001
            print(f"peak measurement failed: {ex}", file=sys.stderr)
        pipe_peak = fp64_peak if args.precision == "f64" else fp32_peak
        achieved_tf = per_gpu_rate * FLOPS_PER_ENV_STEP / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, threads, k_cpu, dt_cpu = cpu_port_rate(args.ref_envs, args.cpu_seconds)
            cpu = {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                   "sample": f"{args.ref_envs} envs x {k_cpu} steps ({dt_cpu:.1f} s) of the same workload, "
                             f"oracle/dockauv_oracle.c with OpenMP on {threads} threads"}
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "envs_per_gpu": N, "envs_total": world * N, "layout": args.layout, "burn_in_steps": args.burn_in,
                       "rollout_steps": args.rollout, "stats_allreduces": n_reduces,
                       "l2": "working set per step ~0.9 GB per GPU >> 126 MB L2, no flush needed"},
            "e2e": {"value": world * N * e2e_steps / e2e_s_max, "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": "env.step_host (dockauv_step_host)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak,
                         "traffic": NCU_TRAFFIC_BYTES_1M_F64 if (N == 1 << 20 and args.precision == "f64") else None,
                         "traffic_source": "profiles/r01/v13_pipeline_ncu_raw.csv (bytes per step = sum of its eight launches)",
                         "peak_source": hbm_src,
                         "bytes_per_env_step": BYTES_PER_ENV_STEP, "kernel_ms": kern_ms,
                         "launches_ms": {n: float(v) for n, v in zip(LAUNCH_NAMES, launch_ms)},
                         "dominant_launch": (LAUNCH_NAMES[int(np.argmax(launch_ms))] if launch_ms.size else "step"),
                         "note": "the path is FP64-pipe-bound (19.7 flop/B vs machine balance ~5.7), see 'pipe'"},
            "pipe": {"bound": "fp64" if args.precision == "f64" else "fp32", "achieved": achieved_tf,
                     "peak": pipe_peak, "unit": "TFLOP/s", "frac": (achieved_tf / pipe_peak) if pipe_peak else None,
                     "flops_per_env_step": FLOPS_PER_ENV_STEP,
                     "executed_pipe_busy_ncu": NCU_FP64_PIPE_BUSY if args.precision == "f64" else None,
                     "note": "frac counts the contract's 17.7 kflop per env-step; the culls skip most ray tests, so the "
                             "pipe itself is ~40 % busy (ncu, per launch above): the launches are latency-bound",
                     "peak_source": "dockauv_measure_peaks FMA micro-kernel on this GPU"},
            "episode_stats": {k: stats[k] for k in ("episodes", "sum_return", "sum_length", "done_collision",
                                                    "done_out_att", "env_steps")},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--layout", default="auto", choices=["auto", "thread_per_env", "warp_rays", "split", "pipeline"])
    ap.add_argument("--burn-in", type=int, default=128)
    ap.add_argument("--rollout", type=int, default=128)
    ap.add_argument("--action-pool", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--ref-envs", type=int, default=16384)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
