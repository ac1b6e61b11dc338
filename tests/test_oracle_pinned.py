"""The CPU restatement (oracle/gne_oracle.cpp) against fixtures produced by the unmodified reference.

tests/golden/*.npz were written by tests/golden/make_golden.py from oracle/_ref (the reference
compiled from /root/reference); the reference itself ships no tests or golden vectors.
Bar: identical (entering, leaving) sequences, bit-equal objective, flows and potentials.
"""
import ctypes as C
import hashlib
import os
import tempfile

import numpy as np
import pytest

from gen_instance import apply_tweak, generate, write_col
from oracle_util import RULES, Instance, Oracle, oracle_lib, trace_md5

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SURVEY_MD5 = {  # SURVEY.md Appendix A (measured on the unmodified reference before this repo existed)
    "LV": "fcdf39b7dc68ed79381ad33527f75d21",
    "QS": "8531c9b53e422e4b6aa0586f82b00f16",
    "CL": "d97d6defef2bbdaded3cf780ac6444fe",
}


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    n, m, seed = (int(v) for v in z["params"])
    mode = str(z["mode"])
    with tempfile.TemporaryDirectory() as tmp:
        col = os.path.join(tmp, "i.col")
        arcs, sup = apply_tweak(*generate(n, m, seed, mode), str(z["tweak"]) if "tweak" in z.files else "")
        write_col(col, n, arcs, sup)
        assert hashlib.md5(open(col, "rb").read()).hexdigest() == str(z["col_md5"]), "generator drifted"
        inst = Instance.from_col(col)
    return z, inst


def check(z, inst, rn, state=True):
    if rn + "_aborted" in z:
        pytest.skip("reference aborts on this instance/rule")
    r1, r2 = (RULES[x] for x in rn.split("+")) if "+" in rn else (RULES[rn], RULES[rn])
    o = Oracle(inst, r1, r2)
    p1, p2 = o.solve_both()
    assert o.status == 0
    e, l = o.trace
    assert (p1, p2) == tuple(int(v) for v in z[rn + "_p"])
    assert np.array_equal(e, z[rn + "_e"]) and np.array_equal(l, z[rn + "_l"])
    assert trace_md5(e, l, p1) == str(z[rn + "_md5"])
    assert o.objective == float(z[rn + "_objective"])
    if rn + "_feasible" in z.files:
        assert o.feasible == bool(z[rn + "_feasible"])
    if state:
        assert np.array_equal(o.flows(), z[rn + "_flows"])
        assert np.array_equal(o.potentials(), z[rn + "_potentials"])
    o.close()


def test_c1_instance_md5_is_the_surveys():
    z, inst = load("c1_wide_s7")
    assert str(z["col_md5"]) == "470d58bff433c7067fab92fe50990e08"
    saved = Instance.load(os.path.join(GOLD, "c1_wide_s7_instance.npz"))
    for k in ("tail", "head", "ub", "cost", "mult", "supply"):
        assert np.array_equal(getattr(saved, k), getattr(inst, k))


@pytest.mark.parametrize("rn", ["LV", "LVW", "QS", "CL"])
def test_c1_trace(rn):
    z, inst = load("c1_wide_s7")
    if rn in SURVEY_MD5:
        assert str(z[rn + "_md5"]) == SURVEY_MD5[rn]
    check(z, inst, rn)


SMALL = ["s_wide_s3", "s_lossy_s4", "s_pow2_s5", "s_near1_s6"]
SMALL_RULES = ["LV", "LVA", "LVE", "FV", "RV", "RWV", "RLV", "RLVW", "CL", "CLW", "SAA", "SAE", "CL2", "QS", "QSW"]


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("rn", SMALL_RULES)
def test_small_all_rules(name, rn):
    z, inst = load(name)
    check(z, inst, rn)


EDGE = ["e_infeasible_s11", "e_nosupply_s12", "e_two_nodes_s13", "e_three_nodes_s14"]
EDGE_RULES = ["LV", "FV", "CL", "CL2", "QS", "RV"]


@pytest.mark.parametrize("name", EDGE)
@pytest.mark.parametrize("rn", EDGE_RULES)
def test_edge_cases(name, rn):
    """Infeasible demand (Phase II is skipped, genneteq.cpp:70-73), zero supplies (every pivot degenerate, objective 0),
    and the smallest graphs the format allows."""
    z, inst = load(name)
    check(z, inst, rn)


MIXED = ["QS+LV", "LV+QS", "CL+LVW", "FV+CL", "CL2+QS", "RV+FV"]


@pytest.mark.parametrize("rn", MIXED)
def test_mixed_phase_rules(rn):
    """Different pivot rules in Phase I and Phase II (`-1 A -2 B`, main.cpp:87-92): cursors, candidate lists and the
    drand48 stream carry over the phase boundary exactly as in the reference."""
    z, inst = load("s_mixed_s8")
    check(z, inst, rn)


def test_edge_fixture_flags():
    assert not bool(np.load(os.path.join(GOLD, "e_infeasible_s11.npz"))["LV_feasible"])
    assert bool(np.load(os.path.join(GOLD, "e_nosupply_s12.npz"))["LV_feasible"])


def test_lossy3k_lv_trace_md5():
    z, inst = load("lossy3k_s5")
    assert str(z["LV_md5"]) == "aec265f38ddd3e1d1999fb257d45d4b5"
    check(z, inst, "LV", state=False)


def test_drand48_restatement_matches_libc():
    lib = oracle_lib()
    libc = C.CDLL(None)
    libc.drand48.restype = C.c_double
    libc.seed48.restype = C.POINTER(C.c_ushort)
    # the reference never seeds: glibc's never-seeded state is X = 0
    zero = (C.c_ushort * 3)(0, 0, 0)
    libc.seed48(zero)
    st = C.c_ulonglong(lib.orc_drand48_initial_state())
    for _ in range(20000):
        assert lib.orc_drand48(C.byref(st)) == libc.drand48()
    first = C.c_ulonglong(0)
    assert lib.orc_drand48(C.byref(first)) == 0xB / 2.0 ** 48


def test_stateless_kernels_agree_with_solver_state():
    """orc_price_* on the extracted nonbasic SoA reproduce what the solver's own rules pick."""
    lib = oracle_lib()
    inst = Instance.load(os.path.join(GOLD, "c1_wide_s7_instance.npz"))
    from oracle_util import dp, ip, lp, bp
    for K in (0, 37, 600, 2500):
        o = Oracle(inst, RULES["LV"])
        o.solve(True, K)
        pi = o.compute_potentials()
        nb = o.nonbasic_soa()
        N = len(nb["order"])
        lst = np.zeros(N, np.int64)
        L = C.c_double(0)
        anyv = C.c_int(0)
        k = lib.orc_price_lv(C.c_long(N), ip(nb["tail"]), ip(nb["head"]), dp(nb["cost"]), dp(nb["mult"]),
                             bp(nb["state"]), dp(pi), lp(lst), C.c_long(N), C.byref(L), C.byref(anyv))
        cand = set(int(v) for v in nb["order"][lst[:k]])
        o.solve(True, 1)               # one more pivot from the same state
        e, _ = o.trace
        if len(e) > K:
            assert int(e[K]) in cand
        o.close()


@pytest.mark.parametrize("name,n,m,seed,rule,K", [("c2_qs", 100_000, 2_000_000, 21, "QS", 2000), ("c2_lv", 100_000, 2_000_000, 21, "LV", 250),
                                                  ("c4_lv", 50_000, 1_000_000, 1000, "LV", 400)])
def test_port_equals_reference_on_large_prefix_windows(name, n, m, seed, rule, K):
    """The port against the unmodified reference BEYOND the small fixtures: prefix windows of the configs[1] / configs[3]
    instances (native generator, 100k nodes / 2M arcs and 50k / 1M) written by the reference itself
    (tests/golden/ref_windows.npz); sized for the CPU suite - the GPU suite compares the full windows."""
    from oracle_util import generate_native, ref_window
    re_, rl_, _ = ref_window(name)
    g = generate_native(n, m, seed, "lossy")
    inst = Instance(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"])
    o = Oracle(inst, RULES[rule], trace_cap=K + 16)
    o.solve(True, K)
    e, l = o.trace
    assert len(e) == K
    assert np.array_equal(e, re_[:K]) and np.array_equal(l, rl_[:K])
    o.close()


def test_native_generator_library_matches_the_product_generator():
    """oracle/libgne_gen.so (what the reference arm of bench.py and make_ref_windows.py load) and the product's
    gne_generate_instance are the same source file with the same flags: identical instances."""
    import genneteq_b200 as gne
    from oracle_util import generate_native
    a = gne.generate_instance(5000, 60000, 77, "lossy")
    b = generate_native(5000, 60000, 77, "lossy")
    for k in ("tail", "head", "ub", "cost", "mult", "supply"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("rn", ["FV", "CL2", "CLW2", "RV", "RLV", "QS"])
def test_medium_fixture_rules_spanning_several_tiles(rn):
    """600 nodes / 9000 arcs (m_wide_s9): first-violator, both candidate-list-2 ids (CLW2 has no other fixture), the random
    rules and quickSelect on a list of several 2048-position tiles; port == unmodified reference."""
    z, inst = load("m_wide_s9")
    check(z, inst, rn)
