#!/usr/bin/env python3
"""Prefix windows of the LARGE workloads straight from the unmodified reference (oracle/_ref/libgne_ref.so).

    python tests/golden/make_ref_windows.py [name ...]         # default: every window below

The big benchmark instances come from the deterministic native generator (genneteq_b200/csrc/gne_generate.cpp,
compiled stand-alone as oracle/libgne_gen.so - no product library is mapped here), are initialised through
ref_driver.cpp's sparse initialiser and solved by the reference's own computePotentials / identifyEnteringVariable /
pivot for the first K pivots of Phase I.  Stored per window in tests/golden/ref_windows.npz: entering / leaving
varIndex per pivot, the md5 of the trace text ("P <iter> <e> <l>\n"), and the generator arguments.

One reference process per window (Tree keeps static pointers, trees.h:83-86).  C3 costs ~1.2 s per pivot.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "ref_windows.npz")
GEN_SO = os.path.join(ROOT, "oracle", "libgne_gen.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgne_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
RULES = dict(LV=0, QS=16)
MODES = dict(wide=0, lossy=1, near1=2, pow2=3)

# name: (nodes, arcs, seed, mode, rule, pivots)
WINDOWS = {
    "c2_lv": (100_000, 2_000_000, 21, "lossy", "LV", 2000),
    "c2_qs": (100_000, 2_000_000, 21, "lossy", "QS", 2000),
    "c3_lv": (1_000_000, 20_000_000, 31, "lossy", "LV", 600),
    "c4_lv": (50_000, 1_000_000, 1000, "lossy", "LV", 2000),
}


def generate(n, m, seed, mode):
    gen = C.CDLL(GEN_SO)
    tail = np.empty(m, np.int32); head = np.empty(m, np.int32)
    ub = np.empty(m); cost = np.empty(m); mult = np.empty(m); supply = np.empty(n)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = gen.gne_generate_instance_impl(C.c_int(n), C.c_int64(m), C.c_uint64(seed), C.c_int(MODES[mode]), p(tail, C.c_int32),
                                        p(head, C.c_int32), p(ub, C.c_double), p(cost, C.c_double), p(mult, C.c_double),
                                        p(supply, C.c_double))
    assert rc == 0, rc
    return dict(n=n, tail=tail, head=head, ub=ub, cost=cost, mult=mult, supply=supply)


def trace_md5(e, l):
    h = hashlib.md5()
    for k in range(len(e)):
        h.update(("P %d %d %d\n" % (k, e[k], l[k])).encode())
    return h.hexdigest()


def child(name, out):
    n, m, seed, mode, rule, K = WINDOWS[name]
    g = generate(n, m, seed, mode)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    orc = C.CDLL(ORACLE_SO)
    orc.orc_big_m.restype = C.c_double
    big_m = orc.orc_big_m(C.c_long(m), dp(g["cost"]), dp(g["ub"]))        # graph.cpp:43 (numArcs x max cost x max capacity)
    ref = C.CDLL(REF_SO)
    ref.ref_solve.restype = C.c_long
    t0 = time.time()
    ref.ref_init_sparse(n, C.c_long(m), ip(g["tail"]), ip(g["head"]), dp(g["ub"]), dp(g["cost"]), dp(g["mult"]), dp(g["supply"]),
                        C.c_double(big_m))
    ref.ref_set_rules(RULES[rule], RULES[rule])
    e = np.zeros(K, np.int32); l = np.zeros(K, np.int32)
    ref.ref_set_trace(ip(e), ip(l), C.c_long(K))
    fin = C.c_int(0)
    ref.ref_solve(1, C.c_long(K), C.byref(fin))
    got = ref.ref_trace_len()
    assert got == K and fin.value == 0, (got, fin.value)
    np.savez(out, e=e, l=l, secs=time.time() - t0)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[3])
    names = sys.argv[1:] or sorted(WINDOWS)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port", "gen", "ref"], stdout=subprocess.DEVNULL)
    data = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        tmp = "/tmp/ref_window_%s.npz" % name
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", name, tmp])
        z = np.load(tmp)
        n, m, seed, mode, rule, K = WINDOWS[name]
        data[name + "_e"] = z["e"]; data[name + "_l"] = z["l"]
        data[name + "_md5"] = np.array(trace_md5(z["e"], z["l"]))
        data[name + "_args"] = np.array(json.dumps(dict(n=n, m=m, seed=seed, mode=mode, rule=rule, pivots=K)))
        print("%s: %d pivots in %.1f s, md5 %s" % (name, K, float(z["secs"]), data[name + "_md5"]))
        np.savez_compressed(OUT, **data)


if __name__ == "__main__":
    main()
