#!/usr/bin/env python
"""Builds tuning variants of the library side by side: gym_dockauv_b200/_lib/libdockauv_<tag>.so with extra nvcc flags.

    python profiles/tools/build_variants.py a3="-DDOCKAUV_MINB_A=3" a5="-DDOCKAUV_MINB_A=5" ...

Each variant is one `python -m gym_dockauv_b200.build --force` with DOCKAUV_LIB_OUT / DOCKAUV_NVCC_EXTRA set; two run
at a time.  Load one with DOCKAUV_LIB=<path> (profiles/tools/ab.sh).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(item):
    tag, flags = item
    env = dict(os.environ)
    env["DOCKAUV_LIB_OUT"] = os.path.join(ROOT, "gym_dockauv_b200", "_lib", f"libdockauv_{tag}.so")
    env["DOCKAUV_NVCC_EXTRA"] = flags
    r = subprocess.run([sys.executable, "-m", "gym_dockauv_b200.build", "--force"], cwd=ROOT, env=env, capture_output=True, text=True)
    return tag, r.returncode, (r.stdout + r.stderr)[-400:]


if __name__ == "__main__":
    items = [a.split("=", 1) for a in sys.argv[1:]]
    with ThreadPoolExecutor(max_workers=2) as ex:
        for tag, rc, log in ex.map(build, items):
            print(tag, "ok" if rc == 0 else "FAILED\n" + log, flush=True)
