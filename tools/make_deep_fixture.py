#!/usr/bin/env python3
"""Record a deep-state basis for `bench.py --window deep` (run on the GPU box).

    python tools/make_deep_fixture.py <workload> <P> <D> <out.npz>

Solves Phase I of the workload for P pivots (call 1), re-enters the solve for D more pivots (call 2: loopIter restarts
and the flows are recomputed exactly as the reference's solve() does on re-entry) and stores both traces.  The file is
then checked against the UNMODIFIED reference with tests/golden/verify_deep_fixture.py (its own pivot() must reproduce
every leaving arc of call 1, and its own solve must continue with exactly the pivots of call 2) before it is committed
under tests/golden/.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genneteq_b200 as gne
import bench

wl, P, D, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
n, m, seed, mode, rule, desc = bench.WORKLOADS[wl]
g = gne.generate_instance(n, m, seed, mode)
s = gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1=rule, trace_cap=P + D + 16)
t0 = time.time()
it, fin = s.solve(True, P)
t1 = time.time()
assert not fin and it == P, (it, fin)
s.solve(True, D)
e, l = s.trace
assert len(e) == P + D
np.savez_compressed(out, e=e[:P], l=l[:P], e2=e[P:], l2=l[P:], workload=wl, n=n, m=m, seed=seed, mode=mode, rule=rule)
print("%s: %d pivots in %.1f s (%.3f ms / pivot), + %d re-entered pivots in %.2f s; saved %s" % (desc, P, t1 - t0, 1e3 * (t1 - t0) / P, D, time.time() - t1, out))
