#!/usr/bin/env python3
"""Tuning harness: time the LV tile kernel and its experimental variants on a resident workload."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genneteq_b200 as gne
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c3-lv"
n, m, seed, mode, rule, desc = bench.WORKLOADS[wl]
g = gne.generate_instance(n, m, seed, mode)
s = gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1="LV")
s.solve(True, 20)
dev = s.device()
alg = 25.0 * m + 8.0 * n
print(desc)
for name, mean, best in dev.bench_variants(20):
    print("%-40s mean %8.2f us  best %8.2f us  %7.1f GB/s (%.1f%% of 6533.8)" % (name, mean * 1e3, best * 1e3, alg / (mean * 1e-3) / 1e9, 100 * alg / (mean * 1e-3) / 1e9 / 6533.8))
