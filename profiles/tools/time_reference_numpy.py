#!/usr/bin/env python
"""Times the UNMODIFIED reference (numpy, /root/reference) on this container's CPU cores -- the C1 configuration of
BASELINE.json and the C4 workload of bench.py -- and records the result as profiles/r02/reference_numpy_cpu.json.

    python profiles/tools/time_reference_numpy.py [--steps 1000] [--procs N]

/root/reference does not exist on the GPU box, so this number cannot be taken there; bench.py prints the recorded
file as `cpu_baseline_numpy` (kind "reference-numpy (recorded)") next to the live C-port baseline.  Timing follows
SURVEY.md 8(d): time.perf_counter around the step loop only, logging / storage / rendering disabled, stdout
suppressed, auto-reset on done, one process per core for the multi-process figure.
"""
import argparse
import contextlib
import datetime
import io
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def run_case(args):
    case, steps, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    import numpy as np
    import ref_shims
    ref_shims.install()
    from make_golden import RADAR64, hook_three_spheres, quiet_config
    from gym_dockauv.envs import docking3d
    np.seterr(all="ignore")
    if case == "C1":
        env = docking3d.SimpleDocking3d(quiet_config())
        hook = None
    else:
        env = docking3d.ObstaclesDocking3d(quiet_config({"radar": RADAR64}))
        hook = hook_three_spheres
    hook_rng = np.random.default_rng(seed + 2)
    rng = np.random.default_rng(seed + 1)
    with contextlib.redirect_stdout(io.StringIO()):
        env.reset(seed=seed)
        if hook:
            hook(env, hook_rng)
        actions = rng.uniform(-1, 1, (steps, 6)).astype(np.float32)
        for a in actions[:20]:
            env.step(a)
        t0 = time.perf_counter()
        episodes = 0
        for a in actions:
            _, _, done, _ = env.step(a)
            if done:
                episodes += 1
                env.reset()
                if hook:
                    hook(env, hook_rng)
        dt = time.perf_counter() - t0
    return steps / dt, episodes


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    args = ap.parse_args()
    out = {"kind": "reference-numpy (recorded)", "unit": "env-steps/s",
           "date": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ"),
           "cpu": cpu_model(), "host_cores": len(os.sched_getaffinity(0)), "steps_per_process": args.steps,
           "reference": "unmodified /root/reference under tests/golden/ref_shims.py (gym / skimage.block_reduce / "
                        "matplotlib shims only), BaseDocking3d.step, docking3d.py:346-402",
           "cases": {}}
    import numpy
    out["numpy"] = numpy.__version__
    for case, label in (("C1", "SimpleDocking3d, BlueROV2, single env, random actions (BASELINE configs[0])"),
                        ("C4", "ObstaclesDocking3d, BlueROV2, 64-ray radar, 5 capsules + 3 spheres (bench.py workload)")):
        one, eps = run_case((case, args.steps, 0))
        with mp.get_context("spawn").Pool(args.procs) as pool:
            t0 = time.perf_counter()
            rates = pool.map(run_case, [(case, args.steps, s) for s in range(args.procs)])
            wall = time.perf_counter() - t0
        out["cases"][case] = {"workload": label, "one_process": {"value": one, "cores": 1, "episodes": eps},
                              "all_cores": {"value": float(sum(r for r, _ in rates)), "cores": args.procs,
                                            "wall_s_including_startup": wall}}
        print(case, out["cases"][case], flush=True)
    path = os.path.join(ROOT, "profiles", "r02", "reference_numpy_cpu.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
