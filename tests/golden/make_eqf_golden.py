#!/usr/bin/env python3
"""Fixtures for instances WITH equal-flow sets: tests/golden/q_*.npz, written by the UNMODIFIED reference.

Run in the build container after ``make -C oracle ref``:   python tests/golden/make_eqf_golden.py

The reference is compiled against oracle/gsl_mini (a restatement of gsl_linalg_LU_decomp / gsl_linalg_LU_solve in GSL 1.x's
operation order; GSL itself is absent from this image), reads each text instance through its own dense loader (graph.cpp)
and solves with its own code; the driver only adds the trace line the reference keeps commented out (gnePivot.cpp:49-50).
Parity of this path is therefore pinned to "reference + restated GSL", not to a GSL build.

An instance = a generated graph (gen_instance.generate) of which `p` disjoint groups of `k` arcs are declared equal-flow
sets (last field of the `a` line, 1-based).  "shuffled" writes the `a` lines in random order: d_r(i) is accumulated in FILE
order (graph.cpp:403-410), variables are numbered in (tail, head) order (gneInit.cpp:82-105).
"""
import hashlib
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from gen_instance import generate  # noqa: E402
from oracle_util import REF_DRIVER, RULES  # noqa: E402


def make_instance(n, m, seed, mode, p, k, shuffled=False, cap_scale=1.0, supply_scale=1.0):
    arcs, sup = generate(n, m, seed, mode)
    sup = [supply_scale * v for v in sup]               # (x 40: demand beyond the capacities, Phase I ends infeasible; x 0: nothing to ship)
    rng = random.Random(seed * 7919 + 13)
    idx = list(range(len(arcs)))
    rng.shuffle(idx)
    eqf = {}
    q = 0
    for r in range(p):
        for _ in range(k):
            eqf[idx[q]] = r + 1
            q += 1
    order = list(range(len(arcs)))
    if shuffled:
        rng.shuffle(order)
    lines = ["p efpgn %d %d %d\n" % (n, len(arcs), p)]
    for a in order:
        (i, j, ub, cost, g) = arcs[a]
        cap = ub * (cap_scale if a in eqf else 1.0)     # (cap_scale: tighter or looser capacities on the member arcs)
        lines.append("a %d %d 0.0 %.6f %.6f %.6f %d\n" % (i + 1, j + 1, cap, cost, g, eqf.get(a, 0)))
    for i in range(n):
        lines.append("n %d %.6f\n" % (i + 1, sup[i]))
    return "".join(lines)


def make_edge_instance(n=40, m=220, seed=31):
    """Loader edge cases in one file (graph.cpp:382-419, gneInit.cpp:64-103): a set of one arc; a set whose arcs share a node
    (d_r(i) = 1 - mult there); a set with a member of capacity 0 (no variable, not counted in UB / cost / numOrigArcs, but its
    +1 / -mult still sit in d_r); a set nobody names (UB = max double, d_r = 0); a set of two antiparallel arcs; a duplicate
    `a` line (the later one overwrites capacity, cost and multiplier; the arc count of the problem line includes it)."""
    arcs, sup = generate(n, m, seed, "wide")
    key = {(a[0], a[1]): k for k, a in enumerate(arcs)}
    eqf = {}
    rng = random.Random(seed * 31 + 7)
    free = [k for k, a in enumerate(arcs) if (a[0] + 1) % n != a[1]]            # keep the ring out of the sets
    rng.shuffle(free)
    take = iter(free)
    for _ in range(3):
        eqf[next(take)] = 1
    eqf[next(take)] = 2
    # set 3: u -> v and v -> w
    by_tail = {}
    for k in free:
        by_tail.setdefault(arcs[k][0], []).append(k)
    for k in free:
        if k in eqf:
            continue
        nxt = [q for q in by_tail.get(arcs[k][1], []) if q not in eqf and q != k]
        if nxt:
            eqf[k] = 3; eqf[nxt[0]] = 3
            break
    zero_cap = None
    for _ in range(3):
        k = next(take)
        while k in eqf:
            k = next(take)
        eqf[k] = 4
        zero_cap = k
    # set 5 stays empty; set 6: an antiparallel pair (added if the generator made none)
    pair = [(k, key[(a[1], a[0])]) for k, a in enumerate(arcs) if (a[1], a[0]) in key and k not in eqf and key[(a[1], a[0])] not in eqf]
    extra = []
    if pair:
        eqf[pair[0][0]] = 6; eqf[pair[0][1]] = 6
    else:
        k = next(take)
        while k in eqf:
            k = next(take)
        eqf[k] = 6
        extra.append((arcs[k][1], arcs[k][0], 37.5, 12.25, 0.9, 6))
    dup = next(take)
    while dup in eqf:
        dup = next(take)
    lines = []
    for k, (i, j, ub, cost, g) in enumerate(arcs):
        cap = 0.0 if k == zero_cap else ub
        lines.append("a %d %d 0.0 %.6f %.6f %.6f %d\n" % (i + 1, j + 1, cap, cost, g, eqf.get(k, 0)))
    for (i, j, ub, cost, g, r) in extra:
        lines.append("a %d %d 0.0 %.6f %.6f %.6f %d\n" % (i + 1, j + 1, ub, cost, g, r))
    (i, j, ub, cost, g) = arcs[dup]
    lines.append("a %d %d 0.0 %.6f %.6f %.6f 0\n" % (i + 1, j + 1, ub * 0.5, cost + 3.0, g))      # the duplicate line
    head = ["p efpgn %d %d %d\n" % (n, len(lines), 6)]
    return "".join(head + lines + ["n %d %.6f\n" % (i + 1, sup[i]) for i in range(n)])


def parse_col(text):
    """(n, p, tail, head, ub, cost, mult, eqf, supply): arcs in FILE order, values as sscanf reads them."""
    n = p = 0
    tail, head, ub, cost, mult, eqf = [], [], [], [], [], []
    sup = None
    for line in text.splitlines():
        if not line.strip():
            continue
        f = line.split()
        if f[0] == "p":
            n, p = int(f[2]), int(f[4])
            sup = [0.0] * n
        elif f[0] == "a":
            tail.append(int(f[1]) - 1); head.append(int(f[2]) - 1); ub.append(float(f[4])); cost.append(float(f[5]))
            mult.append(float(f[6])); eqf.append(int(f[7]))
        elif f[0] == "n":
            sup[int(f[1]) - 1] = float(f[2])
    return (n, p, np.array(tail, np.int32), np.array(head, np.int32), np.array(ub), np.array(cost), np.array(mult),
            np.array(eqf, np.int32), np.array(sup))


def ref_cli(col, rule1, rule2, tmp):
    tr = os.path.join(tmp, "trace.txt")
    dump = os.path.join(tmp, "dump.txt")
    r = subprocess.run([REF_DRIVER, "-1", str(rule1), "-2", str(rule2), "-T", tr, "-D", dump, col],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    if r.returncode != 0:
        return None
    last = [ln for ln in r.stdout.splitlines() if ln.startswith("REF pivots")][-1].split()
    p1, p2 = int(last[2]), int(last[3])
    txt = open(tr).read()
    rows = [ln.split() for ln in txt.splitlines()]
    e = np.array([int(x[2]) for x in rows], np.int32)
    l = np.array([int(x[3]) for x in rows], np.int32)
    flows, pots, obj = [], [], None
    for ln in open(dump):
        f = ln.split()
        if f[0] == "objective":
            obj = float(f[1])
        elif f[0] == "x":
            flows.append(float(f[2]))
        elif f[0] == "pi":
            pots.append(float(f[2]))
    feasible = "Problem is infeasible." not in r.stdout
    return dict(e=e, l=l, p1=p1, p2=p2, objective=obj, flows=np.array(flows), potentials=np.array(pots),
                trace_md5=hashlib.md5(txt.encode()).hexdigest(), feasible=int(feasible))


def make(name, n, m, seed, mode, p, k, rules, shuffled=False, keep_col=False, cap_scale=1.0, text=None, supply_scale=1.0):
    if text is None:
        text = make_instance(n, m, seed, mode, p, k, shuffled, cap_scale, supply_scale)
    out = {}
    (n_, p_, tail, head, ub, cost, mult, eqf, sup) = parse_col(text)
    # the arrays the C ABI takes: arcs sorted by (tail, head) - for a shuffled file d_r(i) is then summed in another
    # order than the file's, which is why the shuffled fixture also carries the text
    o = np.lexsort((head, tail))
    pairs = set(zip(tail.tolist(), head.tolist()))
    out.update(n=n_, p=p_, tail=tail[o], head=head[o], ub=ub[o], cost=cost[o], mult=mult[o], eqf=eqf[o], supply=sup,
               col_md5=hashlib.md5(text.encode()).hexdigest(), shuffled=int(shuffled),
               arrays_ok=int(len(pairs) == len(tail)))     # 0: duplicate `a` lines - only the text describes the instance
    if keep_col:
        out["col_text"] = np.frombuffer(text.encode(), np.uint8)
    nvars = int(np.sum((ub > 0.0) & (eqf == 0))) + n_
    with tempfile.TemporaryDirectory() as tmp:
        col = os.path.join(tmp, "inst.col")
        open(col, "w").write(text)
        for rn in rules:
            r, r2 = (RULES[x] for x in rn.split("+")) if "+" in rn else (RULES[rn], RULES[rn])
            g = ref_cli(col, r, r2, tmp)
            if g is None:
                print("  %s %s: reference aborted" % (name, rn))
                out[rn + "_aborted"] = 1
                continue
            sets_in = int(np.sum(g["e"] >= nvars))
            sets_out = int(np.sum(g["l"] >= nvars))
            print("  %s %-5s pivots %d+%d obj %.9f sets entering %d leaving %d md5 %s" %
                  (name, rn, g["p1"], g["p2"], g["objective"], sets_in, sets_out, g["trace_md5"]))
            out[rn + "_e"] = g["e"]; out[rn + "_l"] = g["l"]
            out[rn + "_p"] = np.array([g["p1"], g["p2"]])
            out[rn + "_objective"] = g["objective"]
            out[rn + "_md5"] = g["trace_md5"]
            out[rn + "_feasible"] = g["feasible"]
            out[rn + "_flows"] = g["flows"]
            out[rn + "_potentials"] = g["potentials"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    allr = ["LV", "LVW", "LVA", "LVE", "FV", "RV", "RWV", "RLV", "RLVW", "CL", "CLW", "SAA", "SAE", "CL2", "CLW2", "QS", "QSW"]
    which = sys.argv[1:] or ["small", "edge", "ends", "medium"]
    if "small" in which:
        make("q_wide_s21", 60, 400, 21, "wide", 4, 3, allr, keep_col=True)
        make("q_lossy_s22", 150, 1200, 22, "lossy", 8, 4, allr)
        make("q_pow2_s23", 100, 800, 23, "pow2", 5, 3, allr)
        make("q_near1_s24", 80, 600, 24, "near1", 6, 2, allr, shuffled=True, keep_col=True)
        make("q_mixed_s25", 90, 700, 25, "wide", 5, 3, ["QS+LV", "LVE+CL", "FV+LVA", "LVW+QSW", "CL2+QS"])
    if "edge" in which:
        # (the arrays of this fixture are the file's lines in (tail, head) order, duplicate included: only the text goes
        #  through the loader the way the reference reads it - the tests use col_text)
        make("q_edge_s31", 0, 0, 0, "", 0, 0, ["LV", "LVE", "QS", "CL", "CL2", "FV"], keep_col=True, text=make_edge_instance())
    if "ends" in which:
        # how a solve with sets ends when it cannot end well: infeasible demand, no supply at all, tight set capacities, and
        # the smallest graphs that can hold a set
        er = ["LV", "LVE", "QS", "CL", "CL2", "FV", "RV"]
        make("q_infeasible_s41", 70, 350, 41, "wide", 4, 3, er, supply_scale=40.0)
        make("q_nosupply_s42", 40, 200, 42, "wide", 3, 2, er, supply_scale=0.0)
        make("q_tightsets_s43", 80, 500, 43, "lossy", 6, 3, er, cap_scale=0.05)
        make("q_three_nodes_s44", 3, 6, 44, "pow2", 1, 2, er)
    if "medium" in which:
        # several 2048-position tiles of nonbasic arcs, more sets.  (At this size the reference often ends with a NaN objective -
        # a singular s x s system, no pivot-size check in its LU use; rule CL does here.  The fixture keeps that run: the
        # restatement has to walk through the same NaNs to the same pivots.)
        make("q_pow2_s29", 600, 9000, 29, "pow2", 10, 4, ["LV", "LVW", "QS", "FV", "CL", "CL2", "CLW2"])
