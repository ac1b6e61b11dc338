#!/usr/bin/env python3
"""Full solve of a generated instance WITH equal-flow sets: the unmodified reference (oracle/_ref/gne_ref_driver, built
against oracle/gsl_mini, 1 thread, its own dense loader) next to the GPU path (gne_solver_create_from_col + gne_solver_solve),
traces compared.  usage: eqf_solve.py [n m seed mode p k rule] [--ref-only] [--budget=K]      (profiles/r2_eqf_solve.txt)
--budget=K stops both after K Phase I pivots (beyond a few hundred nodes the reference's own solves with sets rarely end
well - its periodic self-check reports flow errors of 1e-2 and Phase I ends in NaN - so the comparison is a prefix window)."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from make_eqf_golden import make_instance  # noqa: E402
from oracle_util import REF_DRIVER, RULES  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
ref_only = "--ref-only" in sys.argv
budget = [int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--budget=")]
budget = budget[0] if budget else -1
n, m, seed, mode, p, k, rule = (args + ["3000", "60000", "35", "lossy", "20", "3", "LV"][len(args):])
n, m, seed, p, k = int(n), int(m), int(seed), int(p), int(k)
text = make_instance(n, m, seed, mode, p, k)
with tempfile.TemporaryDirectory() as tmp:
    col = os.path.join(tmp, "inst.col")
    open(col, "w").write(text)
    tr = os.path.join(tmp, "trace.txt")
    t0 = time.time()
    r = subprocess.run([REF_DRIVER, "-1", str(RULES[rule]), "-2", str(RULES[rule]), "-K", str(budget), "-T", tr, col], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    t_ref = time.time() - t0
    last = [ln for ln in r.stdout.splitlines() if ln.startswith("REF pivots")][-1].split()
    p1, p2, obj = int(last[2]), int(last[3]), float(last[-1])
    rows = [ln.split() for ln in open(tr)]
    re_ = np.array([int(x[2]) for x in rows], np.int32); rl = np.array([int(x[3]) for x in rows], np.int32)
    nv = int(max(re_.max(), rl.max()))
    print("instance: %d nodes / %d arcs / %d sets of %d arcs, %s gains, seed %d, rule %s" % (n, m, p, k, mode, seed, rule))
    print("reference (1 thread, wall incl. its dense loader): %.2f s, pivots %d + %d, objective %.9f, %d pivots with a set entering"
          % (t_ref, p1, p2, obj, int(np.sum(re_ >= nv - p + 1))), flush=True)
    if ref_only:
        sys.exit(0)
    import genneteq_b200 as gne
    t0 = time.time()
    s = gne.Solver(col_path=col, rule1=rule)
    t_load = time.time() - t0
    t0 = time.time()
    if budget >= 0:
        q1, fin = s.solve(True, budget)
        q2 = s.solve(False, budget)[0] if fin else 0
    else:
        q1, q2 = s.solve_both()
    t_gpu = time.time() - t0
    e, l = s.trace
    same = (q1, q2) == (p1, p2) and np.array_equal(e, re_) and np.array_equal(l, rl)
    c = s.counters()
    print("GPU path: load %.2f s, solve %.2f s (pricing %.2f, potentials %.2f, pivot %.2f), pivots %d + %d, objective %.9f, %d of them type II pivots"
          % (t_load, t_gpu, c["pricing_s"], c["potentials_s"], c["pivot_s"], q1, q2, s.objective, int(c["trees_s"])))
    print("traces %s; speed-up (solve only) %.1fx, incl. load %.1fx" % ("identical" if same else "DIFFER", t_ref / t_gpu, t_ref / (t_gpu + t_load)))
    sys.exit(0 if same else 1)
