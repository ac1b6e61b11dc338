#!/usr/bin/env python3
"""Per-rank kernel times of a sharded pass, emulated on ONE GPU: the configs[2] list cut to 1/k of its length
(what each of k ranks prices), all n potentials resident, kernels timed back to back (L2-warm, as in the
per-pivot loop where the same shard is re-read every pivot)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genneteq_b200 as gne

n, m, seed = 1_000_000, 20_000_000, 31
g = gne.generate_instance(n, m, seed, "lossy")
rng = np.random.default_rng(5)
pi = rng.uniform(-100, 100, n)
state = np.where(rng.random(m) < 0.3, 2, 1).astype(np.int8)
want = set(sys.argv[1:]) if len(sys.argv) > 1 else None
for k in (1, 2, 4, 8):
    N = (m // k) // 2048 * 2048
    dev = gne.Device(n, N + 16)
    dev.nb_reset(g["tail"][:N], g["head"][:N], g["cost"][:N], g["mult"][:N], state[:N])
    dev.pot_set_all(pi)
    alg = 25.0 * N + 8.0 * n
    print("--- 1/%d of the list: %d arcs, %.1f MB per pass" % (k, N, alg / 1e6))
    for name, mean, best in dev.bench_variants(20):
        if want and not any(w in name for w in want):
            continue
        print("%-58s mean %8.2f us  best %8.2f us  %7.1f GB/s" % (name, mean * 1e3, best * 1e3, alg / (mean * 1e-3) / 1e9))
    dev.close()
