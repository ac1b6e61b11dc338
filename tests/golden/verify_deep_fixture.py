#!/usr/bin/env python3
"""Check a recorded deep-state trace (tools/make_deep_fixture.py) against the UNMODIFIED reference and install it.

    python tests/golden/verify_deep_fixture.py <recorded.npz> [<installed name>]

1. ref_replay (oracle/ref_driver.cpp replayTraced): the reference's own pivot() is applied to the recorded entering
   arcs of call 1 and must pick every recorded leaving arc;
2. from the basis so reached, the reference's own solve (pricing included) must produce exactly the recorded pivots
   of call 2.
Only then is the file copied to tests/golden/deep_<workload>.npz.
"""
import ctypes as C
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_util import REF_SO, RULES, dp, generate_native, ip, oracle_lib  # noqa: E402

src = sys.argv[1]
fx = np.load(src)
n, m, seed, mode, rule = int(fx["n"]), int(fx["m"]), int(fx["seed"]), str(fx["mode"]), str(fx["rule"])
g = generate_native(n, m, seed, mode)
big_m = oracle_lib().orc_big_m(C.c_long(m), dp(g["cost"]), dp(g["ub"]))
ref = C.CDLL(REF_SO)
ref.ref_solve.restype = C.c_long; ref.ref_replay.restype = C.c_long; ref.ref_trace_len.restype = C.c_long
ref.ref_init_sparse(n, C.c_long(m), ip(g["tail"]), ip(g["head"]), dp(g["ub"]), dp(g["cost"]), dp(g["mult"]), dp(g["supply"]), C.c_double(big_m))
ref.ref_set_rules(RULES[rule], RULES[rule])
e = np.ascontiguousarray(fx["e"], np.int32); l = np.ascontiguousarray(fx["l"], np.int32)
t0 = time.time()
mm = ref.ref_replay(1, ip(e), ip(l), C.c_long(len(e)), 1 if rule.startswith("LV") else 0)
print("replay of %d pivots through the reference's pivot(): first mismatch %d, %.1f s" % (len(e), mm, time.time() - t0))
assert mm == -1
D = len(fx["e2"])
te = np.zeros(D, np.int32); tl = np.zeros(D, np.int32)
ref.ref_set_trace(ip(te), ip(tl), C.c_long(D))
fin = C.c_int(0)
t0 = time.time()
ref.ref_solve(1, C.c_long(D), C.byref(fin))
k = ref.ref_trace_len()
ok = k == D and np.array_equal(te, fx["e2"]) and np.array_equal(tl, fx["l2"])
print("reference continued %d pivots in %.1f s (%.1f ms / pivot): %s" % (k, time.time() - t0, 1e3 * (time.time() - t0) / max(k, 1),
                                                                        "equal to the recorded continuation" if ok else "DIFFERENT"))
assert ok
name = sys.argv[2] if len(sys.argv) > 2 else "deep_%s.npz" % str(fx["workload"]).replace("-", "_")
shutil.copyfile(src, os.path.join(HERE, name))
print("installed", os.path.join(HERE, name))
