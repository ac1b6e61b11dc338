#!/usr/bin/env python3
"""Full solves (both phases, to optimality) where BOTH sides finish: wall time of the GPU-driven solve through the
C ABI next to the unmodified reference (oracle/_ref, one host thread) on the same instance, with the pivot traces
compared.  SURVEY 8(d): full-solve speed-ups are claimed only at sizes the reference completes.

usage: python tools/full_solve.py [rule]        (run on the GPU box; needs oracle/_ref built)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import genneteq_b200 as gne
from oracle_util import RULES, Instance, run_reference, have_reference

rule = sys.argv[1] if len(sys.argv) > 1 else "LV"
CASES = [("configs[0]: 1k nodes / 10k arcs, wide, seed 7", 1000, 10000, 7, "wide"),
         ("3k nodes / 60k arcs, lossy, seed 5", 3000, 60000, 5, "lossy")]
if os.environ.get("GNE_FULL_SOLVE_BIG"):               # the reference needs minutes here (31k pivots at ~4.6 ms)
    CASES.append(("6k nodes / 120k arcs, lossy, seed 9", 6000, 120000, 9, "lossy"))
print("rule %s; reference = unmodified sources, 1 thread; GPU = gne_solver_solve (both phases)" % rule)
print("%-46s %9s %12s %12s %9s %12s %s" % ("instance", "pivots", "reference s", "GPU path s", "speed-up", "GPU wall s", "traces"))
for desc, n, m, seed, mode in CASES:
    inst = Instance.generate(n, m, seed, mode)
    s = gne.Solver(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply, rule1=rule)
    t0 = time.perf_counter()
    p1, p2 = s.solve_both()
    gpu_wall = time.perf_counter() - t0             # includes creating the device handle (allocations, stream) at the first solve
    gpu_s = s.counters()["total_s"]                 # the solve loops only - what the reference's timer covers too
    e, l = s.trace
    obj = s.objective
    s.close()
    if have_reference():
        t0 = time.perf_counter()
        z = run_reference(inst, RULES[rule])
        ref_wall = time.perf_counter() - t0
        ref_s = float(z["timers"][3]) if float(z["timers"][3]) > 0 else ref_wall     # the solve loops only, as on the GPU side
        same = np.array_equal(e, z["e"]) and np.array_equal(l, z["l"]) and obj == float(z["objective"])
        print("%-46s %9d %12.3f %12.3f %8.1fx %12.3f %s" % (desc, p1 + p2, ref_s, gpu_s, ref_s / gpu_s, gpu_wall, "identical" if same else "DIFFER"))
    else:
        print("%-46s %9d %12s %12.3f" % (desc, p1 + p2, "n/a", gpu_s))
