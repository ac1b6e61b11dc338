#!/usr/bin/env python3
"""Where the per-pass fixed cost goes: device steps of one / several handles without L2 flush, with and without the
replayed relabel, host wall next to the event times (run on the GPU box)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import genneteq_b200 as gne

def make(n, m, seed, ctas=0):
    g = gne.generate_instance(n, m, seed, "lossy")
    if ctas: os.environ["GNE_STREAM_CTAS"] = str(ctas)
    s = gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1="LV", trace_cap=64)
    s.solve(True, 30)
    return s

for (n, m, name) in ((1000, 10000, "c1"), (50_000, 1_000_000, "c4"), (100_000, 2_000_000, "c2")):
    s = make(n, m, 7)
    d = s.device()
    for wf in (2, 0):
        d.bench_price_lv(20, False, wf)
        t0 = time.perf_counter(); ms_step, ms_k = d.bench_price_lv(200, False, wf); w = time.perf_counter() - t0
        print("%s 1 handle, relabel %s: kernel %.2f us, step %.2f us, host wall per pass %.2f us" % (name, "replayed" if wf else "none", 1e3 * ms_k.mean(), 1e3 * ms_step.mean(), 1e6 * w / 400))
    s.close()
for B in (2, 4, 8, 16, 32):
    ss = [make(50_000, 1_000_000, 1000 + i, ctas=max(1, 296 // B)) for i in range(B)]
    ds = [x.device() for x in ss]
    for d in ds: d.bench_price_lv(10, False)
    res = [None] * B
    def run(i): res[i] = ds[i].bench_price_lv(100, False)
    th = [threading.Thread(target=run, args=(i,)) for i in range(B)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    w = time.perf_counter() - t0
    busy = max(float(r[0].sum() + r[1].sum()) for r in res) * 1e-3
    print("c4 x %2d handles (%d CTAs each): host wall %.1f ms for %d passes = %.2f us per pass overall, busiest stream %.1f ms, kernel %.1f us, aggregate %.1f G arcs/s" % (
        B, max(1, 296 // B), 1e3 * w, B * 200, 1e6 * w / (B * 200), 1e3 * busy, 1e3 * np.mean([r[1].mean() for r in res]), B * 200 * 1e6 / w / 1e9))
    for x in ss: x.close()
