"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(include/genneteq_b200.h via ctypes); the checker is the oracle (oracle/gne_oracle.cpp) and the
golden traces written by the unmodified reference (tests/golden).  Bar: bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_util import RULES, Instance, Oracle, oracle_lib, dp, ip, lp, bp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 1.0e-13


@pytest.fixture(scope="module")
def gne():
    import genneteq_b200
    assert genneteq_b200.cuda_device_count() > 0, "no CUDA device: these tests have no CPU fallback"
    return genneteq_b200


# ---------------------------------------------------------------- stateless kernels
def random_soa(rng, n, N, mode="cont", frac_upper=0.3, scale=1.0):
    tail = np.sort(rng.integers(0, n, N)).astype(np.int32)
    head = rng.integers(0, n, N).astype(np.int32)
    if mode == "ties":                                  # few distinct values => long exact tie lists
        cost = rng.choice([0.0, 1.0, 2.0], N).astype(np.float64)
        mult = rng.choice([0.5, 0.75, 1.25, 1.5, 2.0], N)
        pi = rng.choice([-1.0, 2.0, 0.5], n).astype(np.float64)
    elif mode == "near":                                # values within a few tolerance bands of each other
        cost = 5.0 + rng.integers(0, 9, N) * 0.5e-13
        mult = np.ones(N)
        pi = np.zeros(n)
        pi[: n // 2] = 10.0
    else:
        cost = rng.uniform(-50, 50, N) * scale
        mult = rng.uniform(0.5, 1.5, N)
        pi = rng.uniform(-100, 100, n) * scale
    state = np.where(rng.random(N) < frac_upper, 2, 1).astype(np.int8)
    return tail, head, cost, mult, state, pi


def oracle_lv(tail, head, cost, mult, state, pi):
    lib = oracle_lib()
    N = len(tail)
    lst = np.zeros(max(N, 1), np.int64); L = C.c_double(0); anyv = C.c_int(0)
    k = lib.orc_price_lv(C.c_long(N), ip(tail), ip(head), dp(cost), dp(mult), bp(state), dp(pi), lp(lst), C.c_long(max(N, 1)),
                         C.byref(L), C.byref(anyv))
    return L.value, lst[:k].copy(), bool(anyv.value)


def oracle_qs(tail, head, cost, mult, state, pi, cursor, want):
    lib = oracle_lib()
    bp_ = C.c_long(0); bv = C.c_double(0); nc = C.c_long(0); sc = C.c_long(0); vi = C.c_long(0)
    lib.orc_price_qs(C.c_long(len(tail)), ip(tail), ip(head), dp(cost), dp(mult), bp(state), dp(pi), C.c_long(cursor),
                     C.c_long(want), C.byref(bp_), C.byref(bv), C.byref(nc), C.byref(sc), C.byref(vi))
    return dict(best_pos=bp_.value, best_viol=bv.value, new_cursor=nc.value, scanned=sc.value, violations=vi.value)


def oracle_violators(tail, head, cost, mult, state, pi, thr):
    lib = oracle_lib()
    N = len(tail)
    pos = np.zeros(max(N, 1), np.int64); viol = np.zeros(max(N, 1))
    k = lib.orc_price_violators(C.c_long(N), ip(tail), ip(head), dp(cost), dp(mult), bp(state), dp(pi), C.c_double(thr),
                                lp(pos), dp(viol), C.c_long(max(N, 1)))
    return pos[:k].copy(), viol[:k].copy()


SIZES = [(50, 1), (50, 7), (300, 2047), (300, 2048), (300, 2049), (1000, 10000), (5000, 262144), (20000, 1000003)]


@pytest.mark.parametrize("n,N", SIZES)
@pytest.mark.parametrize("mode", ["cont", "ties", "near"])
@pytest.mark.parametrize("kernel", ["tiles", "stream"])
def test_price_lv_matches_sequential_rule(gne, n, N, mode, kernel):
    """Both LV kernels (one CTA per tile; TMA-streamed persistent) at every size, incl. ragged tails."""
    rng = np.random.default_rng(n * 7 + N)
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, mode)
    d = gne.Device(n, N + 16)
    d.set_lv_kernel(kernel)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    L, lst, anyv = d.price_lv()
    assert d.last_lv_kernel() == kernel
    oL, olst, oany = oracle_lv(tail, head, cost, mult, state, pi)
    assert anyv == oany
    assert L == oL
    assert np.array_equal(lst, olst)
    # two-step form: length + single members (long exact-tie lists never leave the device)
    bL, bcount, bany = d.price_lv_begin()
    assert (bL, bcount, bany) == (oL, len(olst), oany)
    for idx in sorted(set([0, bcount // 3, bcount // 2, bcount - 1])) if bcount else []:
        assert d.price_lv_at(idx) == int(olst[idx])
    # reduced costs themselves, bit for bit
    rc = d.reduced_costs()
    orc = np.zeros(N)
    oracle_lib().orc_reduced_costs(C.c_long(N), ip(tail), ip(head), dp(cost), dp(mult), dp(pi), dp(orc))
    assert np.array_equal(rc, orc)
    d.close()


def test_price_lv_streamed_kernel_is_the_default_for_large_lists_and_falls_back(gne):
    """>= ~2.4M arcs: automatic choice = TMA-streamed kernel.  Each of its threads can contribute one tie-list
    member; longer lists are reported as overflow, the result must still be exact (ordered-compaction fallback) and
    after repeated overflows the handle holds the streamed kernel off."""
    rng = np.random.default_rng(77)
    n, N = 30000, 2_600_000
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, "cont")
    d = gne.Device(n, N + 16)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    L, lst, anyv = d.price_lv()
    assert d.last_lv_kernel() == "stream"
    oL, olst, oany = oracle_lv(tail, head, cost, mult, state, pi)
    assert (L, anyv) == (oL, oany) and np.array_equal(lst, olst)
    # rising values: every thread keeps raising its maximum, still one member per thread at most
    cost2 = 1.0 + np.arange(N) * 1e-3
    d.nb_reset(tail, head, cost2, mult, np.full(N, 2, np.int8))
    d.pot_set_all(np.zeros(n))
    L, lst, anyv = d.price_lv()
    assert L == cost2[-1] and lst.tolist() == [N - 1] and anyv and d.last_lv_kernel() == "stream"
    # one exact tie over the whole list: rc = cost at zero potentials, every arc at its upper bound
    d.nb_reset(tail, head, np.full(N, 5.0), mult, np.full(N, 2, np.int8))
    for _ in range(4):
        L, count, anyv = d.price_lv_begin()
        assert (L, count, anyv) == (5.0, N, True)
        assert d.price_lv_at(N // 3) == N // 3
    assert d.last_lv_kernel() == "tiles"                 # held off after repeated overflows
    d.close()


def test_price_lv_no_violator_and_huge_magnitudes(gne):
    rng = np.random.default_rng(5)
    n, N = 400, 30000
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, "cont")
    d = gne.Device(n, N + 16)
    d.nb_reset(tail, head, np.full(N, 1e9), mult, np.ones(N, np.int8))     # rc > 0 at lower bound: nothing violates
    d.pot_set_all(np.zeros(n))
    L, lst, anyv = d.price_lv()
    assert (L, len(lst), anyv) == (-1.0, 0, False)
    assert d.price_first() == -1
    # bigM-sized values: ulp >> tolerance, ties are exact-equality only
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, "cont", scale=2.0e9)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    L, lst, anyv = d.price_lv()
    oL, olst, oany = oracle_lv(tail, head, cost, mult, state, pi)
    assert (L, anyv) == (oL, oany) and np.array_equal(lst, olst)
    d.close()


@pytest.mark.parametrize("n,N", [(50, 7), (300, 2049), (1000, 10000), (5000, 262144), (20000, 1000003)])
def test_price_qs_matches_scan(gne, n, N):
    rng = np.random.default_rng(N)
    for mode, fu in (("cont", 0.3), ("ties", 0.5), ("cont", 0.999)):
        tail, head, cost, mult, state, pi = random_soa(rng, n, N, mode, frac_upper=fu)
        if fu > 0.9:                                    # almost nothing violates: the scan wraps the whole list
            cost = np.abs(cost) + 500.0
            state = np.ones(N, np.int8); state[rng.integers(0, N, 3)] = 2
        d = gne.Device(n, N + 16)
        d.nb_reset(tail, head, cost, mult, state)
        d.pot_set_all(pi)
        cursors = [0, 1, N // 3, N - 1, N, N + 5] + [int(c) for c in rng.integers(0, N, 4)]
        wants = [1, 2, n, max(1, N // 7), 10 * N]
        for cur in cursors:
            for want in wants:
                got = d.price_qs(cur, want)
                exp = oracle_qs(tail, head, cost, mult, state, pi, cur, want)
                assert got == exp, (mode, cur, want, got, exp)
        d.close()


@pytest.mark.parametrize("n,N", [(50, 7), (300, 2049), (5000, 262144), (20000, 1000003)])
def test_price_violators_and_first(gne, n, N):
    rng = np.random.default_rng(N + 1)
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, "cont")
    d = gne.Device(n, N + 16)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    for thr in (0.0, 10.0, 140.0, 1e9):
        pos, viol = d.price_violators(thr)
        opos, oviol = oracle_violators(tail, head, cost, mult, state, pi, thr)
        assert np.array_equal(pos, opos) and np.array_equal(viol, oviol)
    opos, _ = oracle_violators(tail, head, cost, mult, state, pi, 0.0)
    assert d.price_first() == (int(opos[0]) if len(opos) else -1)
    d.close()


@pytest.mark.parametrize("n,N", [(3, 40), (300, 2049), (5000, 262144)])
def test_price_scan_list_matches_try_to_add_arc(gne, n, N):
    """updateCandidateList2 / tryToAddArc (gnePivotRules.cpp:613-689): at least 5 arcs, then until clMaxSize violators."""
    rng = np.random.default_rng(N + 9)
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, "cont")
    d = gne.Device(n, N + 16)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    allpos, allviol = oracle_violators(tail, head, cost, mult, state, pi, 0.0)
    isv = np.zeros(N, bool); isv[allpos] = True
    vv = np.zeros(N); vv[allpos] = allviol
    for cur in (0, 3, N // 2, N - 2, N + 7):
        for want in (1, 3, n, 10 * N):
            c = 0 if cur >= N else cur
            exp_pos, k, start = [], c, c
            for i in range(5):                          # numArcsToAlwaysCheck
                if isv[k]: exp_pos.append(k)
                k = (k + 1) % N
                if k == start: break
            exhausted = (k == start)
            while len(exp_pos) < want and not exhausted:
                if isv[k]: exp_pos.append(k)
                k = (k + 1) % N
                exhausted = (k == start)
            pos, viol, nc, sc = d.price_scan_list(cur, want)
            assert np.array_equal(pos, np.array(exp_pos, np.int64)), (cur, want)
            assert np.array_equal(viol, vv[pos])
            assert nc == k
    d.close()


def test_nonbasic_mirror_swap_remove_and_push(gne):
    """setNBarc semantics (gnePivot.cpp:527-600) mirrored with gne_nb_write / gne_nb_set_count."""
    rng = np.random.default_rng(11)
    n, N = 200, 5000
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, "cont")
    d = gne.Device(n, N + 16)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    T, H, Cc, Mm, S = [list(a) for a in (tail, head, cost, mult, state)]
    for step in range(300):
        p = int(rng.integers(0, len(T)))
        rec = (T[p], H[p], Cc[p], Mm[p], S[p])
        last = len(T) - 1                               # removeFromL/U: last element moves into the hole
        if p != last:
            d.nb_write(p, int(T[last]), int(H[last]), float(Cc[last]), float(Mm[last]), int(S[last]))
            T[p], H[p], Cc[p], Mm[p], S[p] = T[last], H[last], Cc[last], Mm[last], S[last]
        for a in (T, H, Cc, Mm, S):
            a.pop()
        d.nb_set_count(len(T))
        if step % 3:                                    # insertIntoL/U: push_back, maybe with the other bound
            ns = 3 - rec[4] if step % 2 else rec[4]
            d.nb_write(len(T), int(rec[0]), int(rec[1]), float(rec[2]), float(rec[3]), int(ns))
            T.append(rec[0]); H.append(rec[1]); Cc.append(rec[2]); Mm.append(rec[3]); S.append(ns)
            d.nb_set_count(len(T))
        if step % 25 == 0:
            a = [np.array(x, dt) for x, dt in ((T, np.int32), (H, np.int32), (Cc, np.float64), (Mm, np.float64), (S, np.int8))]
            L, lst, anyv = d.price_lv()
            oL, olst, oany = oracle_lv(a[0], a[1], a[2], a[3], a[4], pi)
            assert (L, anyv) == (oL, oany) and np.array_equal(lst, olst)
    rt, rh, rcst, rm, rs = d.nb_read(0, len(T))
    assert np.array_equal(rt, np.array(T, np.int32)) and np.array_equal(rs, np.array(S, np.int8))
    assert np.array_equal(rcst, np.array(Cc)) and np.array_equal(rm, np.array(Mm)) and np.array_equal(rh, np.array(H, np.int32))
    d.close()


def test_potential_finalize_matches_oracle(gne):
    """pi = a0 + a1 * theta[slot] (gnePotential.cpp:127-135), multiply then add, bit for bit."""
    rng = np.random.default_rng(3)
    n = 100003
    a0 = rng.uniform(-1e3, 1e3, n); a1 = rng.uniform(-3, 3, n)
    slot = rng.integers(0, 977, n).astype(np.int32)
    theta = rng.uniform(-50, 50, 977)
    d = gne.Device(n, 4096)
    d.pot_update_nodes(np.arange(n, dtype=np.int32), a0, a1, slot)
    d.pot_update_thetas(np.arange(977, dtype=np.int32), theta)
    d.pot_finalize()
    got = d.pot_get_all()
    exp = np.zeros(n)
    for s in range(977):                                # per tree, as finalizePotentialsTypeI does
        idx = np.nonzero(slot == s)[0].astype(np.int32)
        oracle_lib().orc_finalize_potentials(C.c_long(len(idx)), ip(idx), dp(np.ascontiguousarray(a0[idx])),
                                             dp(np.ascontiguousarray(a1[idx])), C.c_double(theta[s]), dp(exp))
    assert np.array_equal(got, exp)
    # dirty update of a few nodes and one theta
    idx = rng.integers(0, n, 50).astype(np.int32)
    idx = np.unique(idx)
    a0[idx] += 1.0
    theta[5] = -7.25
    d.pot_update_nodes(idx, a0[idx], a1[idx], slot[idx])
    d.pot_update_thetas(np.array([5], np.int32), np.array([-7.25]))
    d.pot_finalize()
    got = d.pot_get_all()
    exp = a0 + a1 * theta[slot]
    assert np.array_equal(got, exp)
    d.close()


# ---------------------------------------------------------------- solver level
def load_gold(name):
    from test_oracle_pinned import load
    return load(name)


def solve_gpu(gne, inst, rule, max_pivots=-1):
    r1, r2 = rule.split("+") if "+" in rule else (rule, rule)        # "A+B": Phase I rule A, Phase II rule B
    s = gne.Solver(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply, rule1=r1, rule2=r2)
    p1, f1 = s.solve(True, max_pivots)
    p2, f2 = (0, False)
    if f1:
        p2, f2 = s.solve(False, max_pivots)
    return s, p1, p2


@pytest.mark.parametrize("rn", ["LV", "LVW", "QS", "CL"])
def test_c1_pivot_sequence_equals_reference(gne, rn):
    """configs[0]: 1k nodes / 10k arcs; the golden trace was written by the unmodified reference."""
    z, inst = load_gold("c1_wide_s7")
    s, p1, p2 = solve_gpu(gne, inst, rn)
    e, l = s.trace
    assert (p1, p2) == tuple(int(v) for v in z[rn + "_p"])
    assert np.array_equal(e, z[rn + "_e"]) and np.array_equal(l, z[rn + "_l"])
    assert s.objective == float(z[rn + "_objective"])
    assert np.array_equal(s.flows(), z[rn + "_flows"])
    assert np.array_equal(s.potentials(), z[rn + "_potentials"])
    assert s.feasible
    assert s.counters()["launches"] > p1 + p2
    s.close()


SMALL = ["s_wide_s3", "s_lossy_s4", "s_pow2_s5", "s_near1_s6"]
GPU_RULES = ["LV", "LVA", "LVE", "FV", "RV", "RWV", "RLV", "RLVW", "CL", "CLW", "SAA", "SAE", "CL2", "QS", "QSW"]


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("rn", GPU_RULES)
def test_small_instances_every_rule(gne, name, rn):
    z, inst = load_gold(name)
    s, p1, p2 = solve_gpu(gne, inst, rn)
    e, l = s.trace
    assert (p1, p2) == tuple(int(v) for v in z[rn + "_p"])
    assert np.array_equal(e, z[rn + "_e"]) and np.array_equal(l, z[rn + "_l"])
    assert s.objective == float(z[rn + "_objective"])
    assert np.array_equal(s.flows(), z[rn + "_flows"])
    assert np.array_equal(s.potentials(), z[rn + "_potentials"])
    s.close()


@pytest.mark.parametrize("rn", ["QS+LV", "LV+QS", "CL+LVW", "FV+CL", "CL2+QS", "RV+FV"])
def test_mixed_phase_rules(gne, rn):
    """`-1 A -2 B` (main.cpp:87-92): a different rule per phase; trace, flows, potentials, objective equal the reference's."""
    z, inst = load_gold("s_mixed_s8")
    s, p1, p2 = solve_gpu(gne, inst, rn)
    e, l = s.trace
    assert (p1, p2) == tuple(int(v) for v in z[rn + "_p"])
    assert np.array_equal(e, z[rn + "_e"]) and np.array_equal(l, z[rn + "_l"])
    assert s.objective == float(z[rn + "_objective"])
    assert np.array_equal(s.flows(), z[rn + "_flows"])
    assert np.array_equal(s.potentials(), z[rn + "_potentials"])
    s.close()


EDGE = ["e_infeasible_s11", "e_nosupply_s12", "e_two_nodes_s13", "e_three_nodes_s14"]


@pytest.mark.parametrize("name", EDGE)
@pytest.mark.parametrize("rn", ["LV", "FV", "CL", "CL2", "QS", "RV"])
def test_edge_cases_every_rule(gne, name, rn):
    """Infeasible demand (Phase II skipped, feasibility flag), zero supplies (all pivots degenerate, objective 0, long
    exact-tie lists) and the smallest graphs the format allows: trace, flows, potentials, objective and the
    feasibility verdict equal the unmodified reference's."""
    z, inst = load_gold(name)
    s, p1, p2 = solve_gpu(gne, inst, rn)
    e, l = s.trace
    assert (p1, p2) == tuple(int(v) for v in z[rn + "_p"])
    assert np.array_equal(e, z[rn + "_e"]) and np.array_equal(l, z[rn + "_l"])
    assert s.objective == float(z[rn + "_objective"])
    assert s.feasible == bool(z[rn + "_feasible"])
    assert np.array_equal(s.flows(), z[rn + "_flows"])
    assert np.array_equal(s.potentials(), z[rn + "_potentials"])
    s.close()


def test_col_file_loader_and_error_codes(gne, tmp_path):
    z, inst = load_gold("s_wide_s3")
    col = str(tmp_path / "i.col")
    inst.write_col(col)
    s = gne.Solver(col_path=col, rule1="QS")
    p1, p2 = s.solve_both()
    assert np.array_equal(s.trace[0], z["QS_e"]) and np.array_equal(s.trace[1], z["QS_l"])
    s.close()
    bad = str(tmp_path / "bad.col")
    open(bad, "w").write("p efpgn 3 2 1\na 1 2 0.0 5.0 1.0 1.0 1\na 2 3 0.0 5.0 1.0 1.0 0\nn 1 1\nn 2 0\nn 3 -1\n")
    s = gne.Solver(col_path=bad)                         # a file with an equal-flow set loads into the equal-flow engine (tests/test_eqf.py)
    assert s.num_sets == 1 and s.M == 1 + 3 + 1          # one arc outside the set, three artificial self-loops, the set
    s.close()
    open(bad, "w").write("p efpgn 3 2 1\na 1 2 0.0 5.0 1.0 1.0 2\na 2 3 0.0 5.0 1.0 1.0 0\nn 1 1\nn 2 0\nn 3 -1\n")
    with pytest.raises(gne.GneError) as ei:              # names set 2 of 1
        gne.Solver(col_path=bad)
    assert ei.value.code == 2
    with pytest.raises(gne.GneError):
        gne.Solver(col_path=str(tmp_path / "missing.col"))
    with pytest.raises(gne.GneError):                    # rule SD is out of scope
        s = gne.Solver(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply, rule1="SD")
        s.solve(True)


def test_lossy3k_full_solve_equals_reference_trace(gne):
    z, inst = load_gold("lossy3k_s5")
    s, p1, p2 = solve_gpu(gne, inst, "LV")
    e, l = s.trace
    assert (p1, p2) == (3336, 12128)
    assert np.array_equal(e, z["LV_e"]) and np.array_equal(l, z["LV_l"])
    assert s.objective == float(z["LV_objective"])
    s.close()


@pytest.mark.parametrize("rule,K", [("LV", 400), ("QS", 400)])
def test_c2_prefix_window_equals_oracle(gne, rule, K):
    """configs[1]: 100k nodes / 2M arcs (lossy).  First K pivots of Phase I from the initial basis:
    same (entering, leaving) sequence as the oracle, same flows, same potentials."""
    g = gne.generate_instance(100_000, 2_000_000, 21, "lossy")
    inst = Instance(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"])
    s = gne.Solver(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply, rule1=rule)
    s.solve(True, K)
    o = Oracle(inst, RULES[rule])
    o.solve(True, K)
    assert np.array_equal(s.trace[0], o.trace[0]) and np.array_equal(s.trace[1], o.trace[1])
    assert len(s.trace[0]) == K
    assert np.array_equal(s.flows(), o.flows())
    assert np.array_equal(s.potentials(), o.potentials())
    assert s.objective == o.objective
    fs, fo = s.forest(), o.forest()
    for key in ("pred", "predArc", "depth", "thread"):
        assert np.array_equal(fs[key], fo[key])
    s.close(); o.close()


@pytest.mark.parametrize("rn", ["FV", "CL2", "CLW2", "RV", "RLV", "QS"])
def test_medium_fixture_rules_spanning_several_tiles(gne, rn):
    """600 nodes / 9000 arcs: the rules the small fixtures exercise only inside one tile (and CLW2, which has no other
    fixture); trace, flows, potentials, objective equal the unmodified reference's."""
    z, inst = load_gold("m_wide_s9")
    s, p1, p2 = solve_gpu(gne, inst, rn)
    e, l = s.trace
    assert (p1, p2) == tuple(int(v) for v in z[rn + "_p"])
    assert np.array_equal(e, z[rn + "_e"]) and np.array_equal(l, z[rn + "_l"])
    assert s.objective == float(z[rn + "_objective"])
    assert np.array_equal(s.flows(), z[rn + "_flows"])
    assert np.array_equal(s.potentials(), z[rn + "_potentials"])
    s.close()


def test_equal_flow_reduced_costs_match_the_dense_loop(gne):
    """computeEqfRedCost (genneteq.h:321-330): cost - sum_i nodeVals[i] * pi[i], subtracted in node order.  The device walks
    the sparse rows (entries sorted by node); the oracle runs the reference's dense loop over all n nodes: bit-equal, also
    with an empty row, a row touching every node and cancelling terms."""
    rng = np.random.default_rng(17)
    n, p = 5000, 37
    pi = rng.uniform(-100, 100, n)
    dense = np.zeros((p, n))
    for k in range(p):
        cnt = 0 if k == 3 else (n if k == 5 else int(rng.integers(1, 400)))
        idx = np.sort(rng.choice(n, cnt, replace=False))
        dense[k, idx] = rng.choice([1.0, -1.0, 0.5, 2.0, 0.731, -1.25], cnt) * rng.uniform(0.5, 1.5, cnt)
    dense[7, :] = 0.0; dense[7, 10] = 1.0; dense[7, 11] = -1.0; pi[10] = pi[11] = 1e15          # cancellation
    cost = rng.uniform(-50, 50, p)
    exp = np.zeros(p)
    oracle_lib().orc_price_eqf_dense(p, n, dp(np.ascontiguousarray(dense)), dp(cost), dp(pi), dp(exp))
    rows = [np.flatnonzero(dense[k]) for k in range(p)]
    row_start = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    node = np.concatenate(rows).astype(np.int32)
    val = np.concatenate([dense[k, rows[k]] for k in range(p)])
    d = gne.Device(n, 4096)
    d.pot_set_all(pi)
    got = d.price_eqf(row_start, node, val, cost)
    assert np.array_equal(got, exp)
    # the same rows kept on the device, a subset of them named per call (gne_eqf_set_rows / gne_price_eqf_launch / _fetch)
    d.eqf_set_rows(row_start, node, val)
    for ids in (np.arange(p), np.array([5, 3, 36, 7, 0]), np.array([11])):
        assert np.array_equal(d.price_eqf_resident(ids, cost[ids], 0), exp[ids]), ids
    # non-finite potentials (the reference reaches them through a singular s x s system): the dense loop multiplies them by
    # the zero entries too (0 * inf = NaN), so a row that does not touch the bad node must come out NaN all the same, a
    # row whose only bad node carries a non-zero entry keeps its signed infinity
    for bad_nodes, bad_vals in (([123], [np.inf]), ([123, 4000], [np.nan, -np.inf]), ([int(rows[9][0])], [np.inf])):
        pi2 = pi.copy()
        pi2[bad_nodes] = bad_vals
        oracle_lib().orc_price_eqf_dense(p, n, dp(np.ascontiguousarray(dense)), dp(cost), dp(pi2), dp(exp))
        d.pot_set_all(pi2)
        got = d.price_eqf(row_start, node, val, cost)
        assert np.array_equal(got, exp, equal_nan=True), (bad_nodes, got[:12], exp[:12])
        assert np.isnan(got[3]) and np.isnan(exp[3])                # the empty row
        got = d.price_eqf_resident(np.arange(p), cost, len(bad_nodes))           # (the caller states how many potentials are not finite)
        assert np.array_equal(got, exp, equal_nan=True), (bad_nodes, got[:12], exp[:12])
    d.close()


def test_batch_driver_interleaves_without_changing_any_solve(gne):
    """gne_solver_solve_batch (configs[3]: several independent instances on one GPU, fewer host threads than instances, each
    thread stepping its solvers round-robin through launch / ready / complete): every instance must produce exactly the
    pivots, flows and objective it produces alone - for the Dantzig rule (asynchronous passes) and for rules that price
    synchronously (QS, CL) - including instances that finish at different times and the 2- and 3-node graphs."""
    names = ["s_wide_s3", "s_lossy_s4", "s_pow2_s5", "s_near1_s6", "e_two_nodes_s13", "e_three_nodes_s14", "e_nosupply_s12"]
    for rule in ("LV", "QS", "CL"):
        loaded = [load_gold(nm) for nm in names]
        solvers = [gne.Solver(i.n, i.tail, i.head, i.ub, i.cost, i.mult, i.supply, rule1=rule) for (_, i) in loaded]
        r1 = gne.solve_batch(solvers, True, threads=3)
        r2 = gne.solve_batch(solvers, False, threads=2)
        for (z, inst), s, a, b in zip(loaded, solvers, r1, r2):
            assert (a[0], b[0]) == tuple(int(v) for v in z[rule + "_p"])
            e, l = s.trace
            assert np.array_equal(e, z[rule + "_e"]) and np.array_equal(l, z[rule + "_l"])
            assert s.objective == float(z[rule + "_objective"])
            assert np.array_equal(s.flows(), z[rule + "_flows"])
            assert np.array_equal(s.potentials(), z[rule + "_potentials"])
            s.close()


def test_batch_driver_with_launch_groups(gne):
    """gne_solver_solve_batch_grouped: the LV passes of a thread's solvers leave as ONE launch of the batched kernel
    (k_price_lv_batch, up to per_launch instances per grid).  Every instance must still produce exactly the pivots, flows
    and objective it produces alone; rules that price synchronously (QS) pass through untouched."""
    names = ["s_wide_s3", "s_lossy_s4", "s_pow2_s5", "s_near1_s6", "e_two_nodes_s13", "e_three_nodes_s14", "e_nosupply_s12"]
    for rule, threads, per in (("LV", 1, 8), ("LV", 3, 2), ("LV", 2, 3), ("QS", 2, 4)):
        loaded = [load_gold(nm) for nm in names]
        solvers = [gne.Solver(i.n, i.tail, i.head, i.ub, i.cost, i.mult, i.supply, rule1=rule) for (_, i) in loaded]
        st = {}
        # (Phase I with a starting line at pivot 7: instances that end before it - the 2- and 3-node graphs - must not hold the others)
        r1 = gne.solve_batch(solvers, True, threads=threads, per_launch=per, stats=st, align_pivot=7)
        r2 = gne.solve_batch(solvers, False, threads=threads, per_launch=per)
        if rule == "LV":
            total = sum(a[0] for a in r1)
            assert 0 < st["group_launches"] < total, (st, total)          # fewer launches than passes: they did travel together
        for (z, inst), s, a, b in zip(loaded, solvers, r1, r2):
            assert (a[0], b[0]) == tuple(int(v) for v in z[rule + "_p"])
            e, l = s.trace
            assert np.array_equal(e, z[rule + "_e"]) and np.array_equal(l, z[rule + "_l"])
            assert s.objective == float(z[rule + "_objective"])
            assert np.array_equal(s.flows(), z[rule + "_flows"])
            assert np.array_equal(s.potentials(), z[rule + "_potentials"])
            s.close()


def test_lv_group_passes_equal_the_sequential_rule(gne):
    """gne_lv_group: several handles' passes in one grid (lists of one tile, of a few tiles - one CTA each - and lists streamed
    by a share of the CTAs, tie-heavy and near-tie data) give each handle the result of a launch of its own, pass after pass
    (the counters re-arm), whether the group is sent explicitly or launches itself when full."""
    import time
    rng = np.random.default_rng(41)
    shapes = [(50, 7, "cont", 0), (300, 2049, "ties", 0), (5000, 262144, "near", 9), (20000, 700_003, "cont", 37),
              (1000, 10000, "near", 0), (20000, 1_000_003, "ties", 9), (3000, 40_000, "cont", 5)]
    devs, exp = [], []
    for n, N, mode, ctas in shapes:
        tail, head, cost, mult, state, pi = random_soa(rng, n, N, mode)
        d = gne.Device(n, N + 16)
        d.nb_reset(tail, head, cost, mult, state)
        d.pot_set_all(pi)
        if ctas:
            d.set_stream_ctas(ctas)
        devs.append(d)
        oL, olst, oany = oracle_lv(tail, head, cost, mult, state, pi)
        exp.append((oL, len(olst), oany, olst))
    for max_pending in (0, 3):
        g = gne.LvGroup(0, max_pending)
        for rep in range(3):
            for d in devs:
                g.add(d)
            assert g.pending() <= len(devs)
            g.launch()
            assert g.pending() == 0
            for d, (oL, ocount, oany, olst) in zip(devs, exp):
                t0 = time.time()
                while not gne.lib.gne_price_lv_ready(d.h):
                    assert time.time() - t0 < 10.0
                largest = C.c_double(0); count = C.c_int64(0); anyv = C.c_int(0)
                gne._ck(gne.lib.gne_price_lv_complete(d.h, C.byref(largest), C.byref(count), C.byref(anyv)), "gne_price_lv_complete")
                assert (largest.value, count.value, bool(anyv.value)) == (oL, ocount, oany)
                if 0 < ocount <= 4096:
                    pos = C.c_int64(0)
                    for k in (0, ocount - 1):
                        gne._ck(gne.lib.gne_price_lv_at(d.h, k, C.byref(pos)), "gne_price_lv_at")
                        assert pos.value == olst[k]
        # (handles whose tie-heavy lists keep overflowing are held off to the per-tile kernel and launch on their own)
        assert 3 <= g.launch_count() <= (3 if max_pending == 0 else 9)
        g.close()
    # a handle on its own afterwards: the group left its counters armed
    for d, (oL, ocount, oany, olst) in zip(devs, exp):
        assert d.price_lv_begin() == (oL, ocount, oany)
        d.close()


def test_lv_pass_in_two_halves(gne):
    """gne_price_lv_launch / _ready / _complete: the same result as gne_price_lv_begin; ready turns 1 without blocking."""
    import time
    rng = np.random.default_rng(23)
    n, N = 30000, 2_000_003
    tail, head, cost, mult, state, pi = random_soa(rng, n, N)
    d = gne.Device(n, N + 16)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    exp = d.price_lv_begin()
    gne._ck(gne.lib.gne_price_lv_launch(d.h), "gne_price_lv_launch")
    t0 = time.time()
    while not gne.lib.gne_price_lv_ready(d.h):
        assert time.time() - t0 < 10.0
    largest = C.c_double(0); count = C.c_int64(0); anyv = C.c_int(0)
    gne._ck(gne.lib.gne_price_lv_complete(d.h, C.byref(largest), C.byref(count), C.byref(anyv)), "gne_price_lv_complete")
    assert (largest.value, count.value, bool(anyv.value)) == exp
    d.close()


def test_potential_finalize_by_slot(gne):
    """gne_pot_finalize_slots: only the nodes of the named trees get a new pi (a changed theta of a big tree costs one
    sweep over the slot column, no node list); the others keep theirs bit for bit."""
    rng = np.random.default_rng(5)
    n = 70001
    a0 = rng.uniform(-1e3, 1e3, n); a1 = rng.uniform(-3, 3, n)
    slot = rng.integers(0, 40, n).astype(np.int32)
    theta = rng.uniform(-50, 50, 40)
    d = gne.Device(n, 4096)
    d.pot_update_nodes(np.arange(n, dtype=np.int32), a0, a1, slot)
    d.pot_update_thetas(np.arange(40, dtype=np.int32), theta)
    d.pot_finalize()
    before = d.pot_get_all()
    changed = np.array([3, 17, 18, 0, 39, 5, 6, 7, 8, 9, 21], np.int32)          # > 8: two launches
    theta2 = theta.copy(); theta2[changed] = rng.uniform(-50, 50, len(changed))
    d.pot_update_thetas(np.arange(40, dtype=np.int32), theta2)                     # every theta uploaded, only `changed` finalized
    d.pot_finalize_slots(changed)
    got = d.pot_get_all()
    hit = np.isin(slot, changed)
    exp = np.where(hit, a0 + a1 * theta2[slot], before)
    assert np.array_equal(got, exp)
    d.close()


def test_lv_pass_with_a_share_of_the_ctas(gne):
    """gne_dev_set_stream_ctas (batches: every instance's persistent pass gets its share of the CTA slots): the same
    result as with the whole device, for a share of 1, 9 and 37 CTAs."""
    rng = np.random.default_rng(11)
    n, N = 20000, 700_003
    tail, head, cost, mult, state, pi = random_soa(rng, n, N, "near")
    d = gne.Device(n, N + 16)
    d.nb_reset(tail, head, cost, mult, state)
    d.pot_set_all(pi)
    oL, olst, oany = oracle_lv(tail, head, cost, mult, state, pi)
    for ctas in (1, 9, 37, 0):
        d.set_stream_ctas(ctas)
        L, lst, anyv = d.price_lv()
        assert (L, anyv) == (oL, oany) and np.array_equal(lst, olst), ctas
    d.close()


@pytest.mark.parametrize("name,rule", [("c1_wide_s7", "LV"), ("c1_wide_s7", "QS"), ("s_lossy_s4", "LV"), ("s_near1_s6", "CL"), ("lossy3k_s5", "LV")])
def test_fast_math_mode_stays_within_tolerance(gne, name, rule):
    """the FMA-enabled "fast, non-parity" mode (gne_dev_set_fast_math).  The pivot sequence may leave the
    reference's at a tie, but the solve must end at the same optimum: objective within 1e-9 relative of the reference's,
    and - where the reference's final flows / potentials are stored - those within 1e-9 relative too (scale: the largest
    magnitude of the vector).  The switch is per device: it is turned off again and the exact mode must still be bit-exact."""
    z, inst = load_gold(name)
    probe = gne.Device(16, 4096)
    try:
        probe.set_fast_math(True)
        s, p1, p2 = solve_gpu(gne, inst, rule)
        obj = float(z[rule + "_objective"])
        assert abs(s.objective - obj) <= 1e-9 * max(1.0, abs(obj)), (s.objective, obj)
        if rule + "_flows" in z.files:
            fl, rf = s.flows(), z[rule + "_flows"]
            assert np.max(np.abs(fl - rf)) <= 1e-9 * max(1.0, np.max(np.abs(rf)))
            po, rp = s.potentials(), z[rule + "_potentials"]
            assert np.max(np.abs(po - rp)) <= 1e-9 * max(1.0, np.max(np.abs(rp)))
        same = np.array_equal(s.trace[0], z[rule + "_e"])
        print("fast math, %s %s: %d+%d pivots (reference %s), trace %s" % (name, rule, p1, p2, tuple(int(v) for v in z[rule + "_p"]),
                                                                          "identical" if same else "differs"))
        s.close()
    finally:
        probe.set_fast_math(False)
    s, p1, p2 = solve_gpu(gne, inst, rule)
    assert np.array_equal(s.trace[0], z[rule + "_e"]) and np.array_equal(s.trace[1], z[rule + "_l"])
    s.close(); probe.close()


REF_WINDOWS = [("c2_lv", 100_000, 2_000_000, 21, "LV", 2000), ("c2_qs", 100_000, 2_000_000, 21, "QS", 2000),
               ("c4_lv", 50_000, 1_000_000, 1000, "LV", 2000), ("c3_lv", 1_000_000, 20_000_000, 31, "LV", 600)]


@pytest.mark.parametrize("name,n,m,seed,rule,K", REF_WINDOWS)
def test_large_prefix_windows_equal_the_reference_itself(gne, name, n, m, seed, rule, K):
    """configs[1] (LV and the block rule QS, 2000 pivots), configs[3]'s instance 0 (2000 pivots) and configs[2]
    (1M nodes / 20M arcs, 600 pivots): the (entering, leaving) sequence from the initial basis equals, pivot for
    pivot, the one the UNMODIFIED reference produced for the same generated instance
    (tests/golden/ref_windows.npz, written by tests/golden/make_ref_windows.py through oracle/_ref/libgne_ref.so)."""
    from oracle_util import ref_window, REF_WINDOWS as path
    import os
    re_, rl_, md5 = ref_window(name)
    assert len(re_) == K
    g = gne.generate_instance(n, m, seed, "lossy")
    s = gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1=rule, trace_cap=K + 16)
    s.solve(True, K)
    e, l = s.trace
    assert len(e) == K
    bad = np.flatnonzero((e != re_) | (l != rl_))
    assert bad.size == 0, "first differing pivot %d: (%d, %d) vs reference (%d, %d)" % (bad[0], e[bad[0]], l[bad[0]], re_[bad[0]], rl_[bad[0]])
    s.close()


def test_c2_deep_same_state_ab(gne):
    """same-state A/B deep into a GPU-driven solve of configs[1] (100k nodes / 2M arcs).  At
    sampled iterations the state the next pricing pass will read (nonbasic mirror in setNBarc order + potentials)
    is read back from the device and priced by the oracle's scalar pass: largest violation, the whole tie list and
    the quickSelect block must be identical, the next entering variable must be a member of the oracle's list, and
    the extra probe must not disturb the pivot sequence (second solver without probes gives the same trace)."""
    import time
    g = gne.generate_instance(100_000, 2_000_000, 21, "lossy")
    n, m = g["n"], len(g["tail"])
    key = g["tail"].astype(np.int64) * n + g["head"]
    assert np.all(np.diff(key) > 0)
    s = gne.Solver(n, g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1="LV", trace_cap=60000)
    plain = gne.Solver(n, g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], rule1="LV", trace_cap=60000)
    done = 0
    for target in (1000, 12000, 25000, 40000):
        # (re-entering solve() restarts loopIter and recomputes the flows exactly as the reference's solve() would,
        #  so the probe-free solver below is driven through the same sequence of calls)
        s.solve(True, target - done)
        plain.solve(True, target - done)
        done = target
        s.compute_potentials()
        dev = s.device()
        st = s.states()
        N = int(np.count_nonzero(st != 0))                          # |setNBarc|: L_var = 1, U_var = 2, basic = 0
        t, h, c, mu, sta = dev.nb_read(0, N)
        pi = dev.pot_get_all()[:n]
        assert np.array_equal(pi, s.potentials())
        t0 = time.perf_counter()
        L, lst, anyv = oracle_lv(t, h, c, mu, sta, pi)
        cpu_s = time.perf_counter() - t0
        gL, gcount, gany = dev.price_lv_begin()
        kern_ms = dev.last_kernel_ms()
        assert (gL, gcount, gany) == (L, len(lst), anyv)
        for idx in {0, gcount // 2, gcount - 1}:
            assert dev.price_lv_at(idx) == lst[idx]
        for cursor in (0, N // 3, N - 5):
            assert dev.price_qs(cursor, n) == oracle_qs(t, h, c, mu, sta, pi, cursor, n)
        # the variable the solver enters next is a member of the oracle's tie list at this state
        var = np.where(t == h, m + t.astype(np.int64), np.searchsorted(key, t.astype(np.int64) * n + h))
        s.solve(True, 1)
        plain.solve(True, 1)
        done += 1
        assert int(s.trace[0][done - 1]) in set(var[lst].tolist())
        print("pivot %d: |NB| %d  L %.6g  ties %d  oracle pass %.1f ms  device kernel %.4f ms" %
              (target, N, L, len(lst), 1e3 * cpu_s, kern_ms))
    assert len(s.trace[0]) == done
    assert np.array_equal(plain.trace[0], s.trace[0]) and np.array_equal(plain.trace[1], s.trace[1])
    assert np.array_equal(plain.flows(), s.flows())
    s.close(); plain.close()


def test_c3_full_size_pricing_pass_equals_oracle(gne):
    """configs[2] size: 1M nodes / 20M arcs.  One Dantzig pass at the initial Phase-I state and after
    200 pivots, against the oracle's scalar pass over the same 20M-arc structure-of-arrays."""
    g = gne.generate_instance(1_000_000, 20_000_000, 31, "lossy")
    inst = Instance(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"])
    s = gne.Solver(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply, rule1="LV")
    o = Oracle(inst, RULES["LV"])
    for K in (1, 200):
        s.solve(True, K)
        o.solve(True, K)
        assert np.array_equal(s.trace[0], o.trace[0]) and np.array_equal(s.trace[1], o.trace[1])
        pi = o.compute_potentials()
        nb = o.nonbasic_soa()
        dev = s.device()
        # the device's mirror of setNBarc equals the oracle's list, record for record
        rt, rh, rcst, rm, rs = dev.nb_read(0, inst.m)
        assert np.array_equal(rt, nb["tail"]) and np.array_equal(rh, nb["head"]) and np.array_equal(rs, nb["state"])
        assert np.array_equal(rcst, nb["cost"]) and np.array_equal(rm, nb["mult"])
    s.close(); o.close()


def test_c5_full_size_kernels_equal_oracle(gne):
    """configs[4] size: 200k nodes / 100M arcs (2.5 GB of arc columns).  Initial Phase-I state (cost 0 on real arcs,
    pi = 1/(1-mu) of each node's artificial self-loop): both LV kernels, quickSelect from three cursors and the
    first-violator pass against the oracle's scalar passes over the same structure-of-arrays."""
    g = gne.generate_instance(200_000, 100_000_000, 51, "lossy")
    n, N = g["n"], len(g["tail"])
    pi = np.where(g["supply"] >= 0, 1.0 / (1.0 - 0.5), 1.0 / (1.0 - 2.0))
    cost = np.zeros(N)
    state = np.ones(N, np.int8)
    d = gne.Device(n, N + 16)
    d.nb_reset(g["tail"], g["head"], cost, g["mult"], state)
    d.pot_set_all(pi)
    oL, olst, oany = oracle_lv(g["tail"], g["head"], cost, g["mult"], state, pi)
    for kernel in ("stream", "tiles"):
        d.set_lv_kernel(kernel)
        L, lst, anyv = d.price_lv()
        assert (L, anyv) == (oL, oany) and np.array_equal(lst, olst), kernel
    for cur in (0, 41_234_567, N - 3):
        assert d.price_qs(cur, n) == oracle_qs(g["tail"], g["head"], cost, g["mult"], state, pi, cur, n)
    opos, _ = oracle_violators(g["tail"], g["head"], cost, g["mult"], state, pi, 2.9)
    pos, _ = d.price_violators(2.9, cap=len(opos) + 16)
    assert np.array_equal(pos, opos)
    d.close()
