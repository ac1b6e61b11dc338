#!/usr/bin/env python3
"""Synthetic generalized min-cost-flow instance generator.

Writes the reference's text format (graph.cpp:358-447): ``p efpgn n m 0`` / ``a t h LB UB cost
mult eqf`` / ``n id supply``; node ids are 1-based, every node gets an ``n`` line, no
equal-flow sets.  The draw order (Python ``random.Random(seed)``) is part of the fixture
definition: the md5 of ``gen_instance.py 1000 10000 7 out wide`` is pinned in
tests/golden/README.md.

usage: gen_instance.py n m seed out.col mode      mode in {wide, lossy, near1, pow2}
"""
import random
import sys


def generate(n, m, seed, mode):
    """Return (arcs, supply) with arcs = sorted list of (tail, head, ub, cost, gain), 0-based."""
    rng = random.Random(seed)
    pairs = set()
    for i in range(n):
        pairs.add((i, (i + 1) % n))
    while len(pairs) < m:
        i = rng.randrange(n)
        j = rng.randrange(n)
        if i != j:
            pairs.add((i, j))
    pairs = sorted(pairs)
    sup = [0.0] * n
    k = max(1, n // 20)
    nodes = list(range(n))
    rng.shuffle(nodes)
    for i in nodes[:k]:
        sup[i] = float(rng.randint(1, 20)) * (3 if mode == 'lossy' else 1)
    for i in nodes[k:2 * k]:
        sup[i] = -float(rng.randint(1, 20))

    def gain():
        if mode == 'lossy':
            return rng.uniform(0.70, 0.99)
        if mode == 'near1':
            return rng.uniform(0.9, 1.1)
        if mode == 'pow2':
            return rng.choice([0.5, 0.75, 1.25, 1.5, 2.0])
        return rng.uniform(0.5, 1.5)

    arcs = []
    for (i, j) in pairs:  # per-arc draw order: UB, cost, gain
        ub = rng.uniform(10, 100)
        cost = rng.uniform(1, 100)
        g = gain()
        arcs.append((i, j, ub, cost, g))
    return arcs, sup


def apply_tweak(arcs, sup, tweak):
    """Edge-case derivations of a generated instance (tests/golden/e_*.npz): same arcs, altered supplies."""
    if tweak in ("", None):
        return arcs, sup
    if tweak == "supplyx40":                            # demand far beyond the arc capacities: Phase I ends infeasible
        return arcs, [40.0 * v for v in sup]
    if tweak == "nosupply":                             # nothing to ship: both phases end without a pivot
        return arcs, [0.0 for _ in sup]
    raise ValueError("unknown tweak %r" % tweak)


def write_col(path, n, arcs, sup):
    with open(path, 'w') as f:
        f.write("p efpgn %d %d 0\n" % (n, len(arcs)))
        for (i, j, ub, cost, g) in arcs:
            f.write("a %d %d 0.0 %.6f %.6f %.6f 0\n" % (i + 1, j + 1, ub, cost, g))
        for i in range(n):
            f.write("n %d %.6f\n" % (i + 1, sup[i]))


def read_col(path):
    """Parse the text format the way graph.cpp does (LB is read and dropped, graph.cpp:386-410).

    Returns dict(n, tail, head, ub, cost, mult, supply) with arcs sorted by (tail, head) -
    the order GenNetEq::initializeVars assigns varIndex in (gneInit.cpp:82-107).
    """
    n = 0
    arcs = []
    sup = None
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            c = line[0]
            if c == 'p':
                parts = line.split()
                n = int(parts[2])
                sup = [0.0] * n
            elif c in 'ae':
                p = line.split()
                arcs.append((int(p[1]) - 1, int(p[2]) - 1, float(p[4]), float(p[5]), float(p[6])))
            elif c == 'n':
                p = line.split()
                sup[int(p[1]) - 1] = float(p[2])
    arcs.sort(key=lambda a: (a[0], a[1]))
    return dict(n=n, tail=[a[0] for a in arcs], head=[a[1] for a in arcs], ub=[a[2] for a in arcs],
                cost=[a[3] for a in arcs], mult=[a[4] for a in arcs], supply=sup)


if __name__ == '__main__':
    n = int(sys.argv[1]); m = int(sys.argv[2]); seed = int(sys.argv[3]); out = sys.argv[4]; mode = sys.argv[5]
    arcs, sup = generate(n, m, seed, mode)
    write_col(out, n, arcs, sup)
