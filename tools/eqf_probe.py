#!/usr/bin/env python3
"""Debug aid for the equal-flow engine: solve one q_* fixture with one rule under a pivot budget and report where the
trace first leaves the reference's (tests/golden/q_*.npz).  usage: eqf_probe.py fixture rule [budget]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genneteq_b200 as gne  # noqa: E402

name, rn = sys.argv[1], sys.argv[2]
z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
r1, r2 = rn.split("+") if "+" in rn else (rn, rn)
s = gne.Solver(int(z["n"]), z["tail"], z["head"], z["ub"], z["cost"], z["mult"], z["supply"], rule1=r1, rule2=r2,
               eqf=z["eqf"], num_sets=int(z["p"]))
p1, p2 = (int(v) for v in z[rn + "_p"])
budget = int(sys.argv[3]) if len(sys.argv) > 3 else p1 + 50
t0 = time.time()
it, fin = s.solve(True, budget)
e, l = s.trace
k = min(len(e), p1)
bad = np.nonzero((e[:k] != z[rn + "_e"][:k]) | (l[:k] != z[rn + "_l"][:k]))[0]
print("%s %s: phase I %d pivots (reference %d) finished %s in %.2f s; first difference %s; objective %r" %
      (name, rn, it, p1, fin, time.time() - t0, int(bad[0]) if len(bad) else None, s.objective), flush=True)
if len(bad):
    b = int(bad[0])
    print("  at %d: got (%d, %d), reference (%d, %d)" % (b, e[b], l[b], z[rn + "_e"][b], z[rn + "_l"][b]))
if fin and not len(bad):
    t0 = time.time()
    it2, fin2 = s.solve(False, p2 + 50)
    e, l = s.trace
    k = min(len(e), p1 + p2)
    bad = np.nonzero((e[:k] != z[rn + "_e"][:k]) | (l[:k] != z[rn + "_l"][:k]))[0]
    print("  phase II %d pivots (reference %d) finished %s in %.2f s; first difference %s; objective %r (reference %r)" %
          (it2, p2, fin2, time.time() - t0, int(bad[0]) if len(bad) else None, s.objective, float(z[rn + "_objective"])), flush=True)
