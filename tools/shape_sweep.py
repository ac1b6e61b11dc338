#!/usr/bin/env python3
"""Tuning run: the streamed LV pass in each instantiated shape (GNE_ST_SHAPE="stages,CTAs per SM") through bench.py.
usage: python tools/shape_sweep.py [workload] [steps]   (prints one line per shape; run on the GPU box)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else "c3-lv"
steps = sys.argv[2] if len(sys.argv) > 2 else "20"
for shape in ("3,2", "2,2", "2,3", "4,2", "3,1"):
    env = dict(os.environ, GNE_ST_SHAPE=shape)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", steps, "--warmup", "5", "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print("%s shape %s: kernel %.2f us (%.1f%% of %s GB/s), device step %.2f us, e2e %.2f us / pivot, parity %s" % (
            wl, shape, 1e3 * d["roofline"]["kernel_ms"], 100 * d["roofline"]["frac"], d["roofline"]["peak"], 1e3 * d["ms_per_step"],
            1e3 * d["e2e"]["ms_per_pivot"], d["parity"]["vs_ref"]))
    except Exception as ex:
        print("%s shape %s: failed (%r) %s" % (wl, shape, ex, r.stderr[-400:]))
    sys.stdout.flush()
