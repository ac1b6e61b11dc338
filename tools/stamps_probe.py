#!/usr/bin/env python3
"""Where the time of one streamed LV pass goes (globaltimer stamps of a -DGNE_STAMPS build; run on the GPU box):
    make -C genneteq_b200/csrc EXTRA=-DGNE_STAMPS OUT=../lib/libgenneteq_b200_stamps.so OBJ=../lib/obj_stamps
    GNE_LIB=libgenneteq_b200_stamps.so python tools/stamps_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genneteq_b200 as gne

for (n, m, name) in ((1000, 10000, "c1"), (50_000, 1_000_000, "c4"), (100_000, 2_000_000, "c2"), (1_000_000, 20_000_000, "c3")):
    g = gne.generate_instance(n, m, 7, "lossy")
    s = gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1="LV", trace_cap=64)
    s.solve(True, 30)
    d = s.device()
    for wf in (2, 0):
        sys.stderr.write("== %s, relabel %s\n" % (name, "replayed" if wf else "none")); sys.stderr.flush()
        d.bench_price_lv(110, False, wf)
    s.close()
