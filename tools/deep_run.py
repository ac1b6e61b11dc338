#!/usr/bin/env python3
"""Per-window cost of the GPU-driven solve as the basis trees grow (host forest work vs pricing)."""
import sys, os
os.environ.setdefault("GNE_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genneteq_b200 as gne
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c2-lv"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
n, m, seed, mode, rule, desc = bench.WORKLOADS[wl]
g = gne.generate_instance(n, m, seed, mode)
save = sys.argv[4] if len(sys.argv) > 4 else None          # npz path: keep the (entering, leaving) trace of the run
s = gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1=rule, trace_cap=(total + 16) if save else 16)
s.set_progress(every)
it, fin = s.solve(True, total)
P = s.progress()
names = gne.Solver.COUNTERS
print(desc, "| pivots", it, "finished", fin)
print("%8s %9s %9s %9s %9s %9s %9s %9s %10s %8s %10s %10s" % ("pivot", "ms/pivot", "pricing", "potent.", "pivot", "trees", "flows", "kern ms", "maxtree", "launch", "h2d B", "d2h B"))
for i in range(1, len(P)):
    d = P[i] - P[i - 1]
    c = dict(zip(names, d[1:13]))
    k = c["pivots"]
    print("%8d %9.4f %9.4f %9.4f %9.4f %9.4f %9.4f %9.4f %10d %8.2f %10.0f %10.0f" % (P[i][10], 1e3 * d[0] / k, 1e3 * c["pricing_s"] / k, 1e3 * c["potentials_s"] / k,
          1e3 * c["pivot_s"] / k, 1e3 * c["trees_s"] / k, 1e3 * c["flows_s"] / k, c["kernel_ms"] / k, P[i][13], c["launches"] / k, c["h2d_bytes"] / k, c["d2h_bytes"] / k))
if save:
    e, l = s.trace
    np.savez_compressed(save, e=e, l=l, workload=wl, n=n, m=m, seed=seed, mode=mode, rule=rule)
    print("trace of %d pivots saved to %s" % (len(e), save))
