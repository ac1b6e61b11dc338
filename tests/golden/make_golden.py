#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the UNMODIFIED reference (needs /root/reference).

Run in the build container after ``make -C oracle ref``:   python tests/golden/make_golden.py

Every fixture is produced by oracle/_ref/gne_ref_driver reading the text instance through the
reference's own loader (graph.cpp) and solving with the reference's own compiled code; the
driver only adds the trace line the reference keeps commented out (gnePivot.cpp:49-50).
The same instances are also pushed through the driver's sparse initialiser and must give the
identical trace (that is what lets the oracle run sizes the dense loader cannot hold).
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from gen_instance import apply_tweak, generate, write_col  # noqa: E402
from oracle_util import REF_DRIVER, RULES, Instance, run_reference  # noqa: E402


def ref_cli(col, rule1, rule2, tmp):
    tr = os.path.join(tmp, "trace.txt")
    dump = os.path.join(tmp, "dump.txt")
    r = subprocess.run([REF_DRIVER, "-1", str(rule1), "-2", str(rule2), "-T", tr, "-D", dump, col],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
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
        p = ln.split()
        if p[0] == "objective":
            obj = float(p[1])
        elif p[0] == "x":
            flows.append(float(p[2]))
        elif p[0] == "pi":
            pots.append(float(p[2]))
    return dict(e=e, l=l, p1=p1, p2=p2, objective=obj, flows=np.array(flows), potentials=np.array(pots),
                trace_md5=hashlib.md5(txt.encode()).hexdigest())


def make(name, n, m, seed, mode, rules, keep_state=True, save_instance=False, tweak=""):
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        col = os.path.join(tmp, "inst.col")
        arcs, sup = apply_tweak(*generate(n, m, seed, mode), tweak)
        write_col(col, n, arcs, sup)
        out["tweak"] = tweak
        out["col_md5"] = hashlib.md5(open(col, "rb").read()).hexdigest()
        out["params"] = np.array([n, m, seed])
        out["mode"] = mode
        inst = Instance.from_col(col)
        if save_instance:
            inst.save(os.path.join(HERE, name + "_instance.npz"))
        for rn in rules:
            r, r2 = (RULES[x] for x in rn.split("+")) if "+" in rn else (RULES[rn], RULES[rn])    # "A+B": Phase I rule A, Phase II rule B
            g = ref_cli(col, r, r2, tmp)
            if g is None:
                print("  %s %s: reference aborted" % (name, rn))
                out[rn + "_aborted"] = 1
                continue
            z = run_reference(inst, r, r2)
            same = np.array_equal(z["e"], g["e"]) and np.array_equal(z["l"], g["l"]) and z["objective"] == g["objective"]
            assert same, "sparse initialiser diverges from the dense loader on %s %s" % (name, rn)
            print("  %s %-4s pivots %d+%d obj %.9f trace md5 %s" % (name, rn, g["p1"], g["p2"], g["objective"], g["trace_md5"]))
            out[rn + "_e"] = g["e"]
            out[rn + "_l"] = g["l"]
            out[rn + "_p"] = np.array([g["p1"], g["p2"]])
            out[rn + "_objective"] = g["objective"]
            out[rn + "_md5"] = g["trace_md5"]
            out[rn + "_feasible"] = int(z["feasible"])
            if keep_state:
                out[rn + "_flows"] = g["flows"]
                out[rn + "_potentials"] = g["potentials"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "small", "lossy3k"]
    if "c1" in which:
        make("c1_wide_s7", 1000, 10000, 7, "wide", ["LV", "LVW", "QS", "CL"], save_instance=True)
    if "small" in which:
        allr = ["LV", "LVA", "LVE", "FV", "RV", "RWV", "RLV", "RLVW", "CL", "CLW", "SAA", "SAE", "CL2", "QS", "QSW"]
        make("s_wide_s3", 120, 1200, 3, "wide", allr)
        make("s_lossy_s4", 150, 1500, 4, "lossy", allr)
        make("s_pow2_s5", 100, 900, 5, "pow2", allr)
        make("s_near1_s6", 130, 1100, 6, "near1", allr)
    if "mixed" in which:
        make("s_mixed_s8", 140, 1300, 8, "wide", ["QS+LV", "LV+QS", "CL+LVW", "FV+CL", "CL2+QS", "RV+FV"])
    if "edge" in which:
        er = ["LV", "FV", "CL", "CL2", "QS", "RV"]
        make("e_infeasible_s11", 80, 400, 11, "wide", er, tweak="supplyx40")
        make("e_nosupply_s12", 40, 200, 12, "wide", er, tweak="nosupply")
        make("e_two_nodes_s13", 2, 2, 13, "wide", er)
        make("e_three_nodes_s14", 3, 6, 14, "pow2", er)
    if "medium" in which:
        # several 2048-position tiles (so that a list sharded over 2-4 ranks has more than one non-empty shard) for the rules
        # the small fixtures cover only within one tile: FV, CL2, CLW2 (which no other fixture holds), RV, RLV
        make("m_wide_s9", 600, 9000, 9, "wide", ["FV", "CL2", "CLW2", "RV", "RLV", "QS"], save_instance=True)
    if "lossy3k" in which:
        make("lossy3k_s5", 3000, 60000, 5, "lossy", ["LV"], keep_state=False)
