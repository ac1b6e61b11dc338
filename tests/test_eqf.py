"""Equal-flow sets: the equal-flow engine against fixtures written by the unmodified reference
(tests/golden/q_*.npz, make_eqf_golden.py; the reference compiled against oracle/gsl_mini's restatement of GSL's LU).

CPU part: the host logic - Type II trees, tree ids, pivotTII, the (s+1)-vector flows, both s x s systems, potentials -
replayed without a device (entering variables from the trace), every leaving variable and the final state compared.
GPU part: the whole solve with arcs and sets priced on the device, every rule the engine offers, the text loader."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DRAWS = dict(LV=1, LVW=1, LVA=1, LVE=1, FV=0, RV=1, RWV=1, RLV=1, RLVW=1, CL=0, CLW=0, SAA=2, SAE=2, QS=0, QSW=0)
FIXTURES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "q_*.npz")))


@pytest.fixture(scope="module")
def gne():
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import genneteq_b200
    return genneteq_b200


def rules_of(z):
    return sorted(k[:-2] for k in z.files if k.endswith("_e"))


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def same_scalar(a, b):
    return a == b or (np.isnan(a) and np.isnan(b))


def test_fixtures_present():
    assert len(FIXTURES) >= 6, FIXTURES


def arrays_ok(z):
    return "arrays_ok" not in z.files or int(z["arrays_ok"]) == 1


@pytest.mark.parametrize("name", FIXTURES)
def test_eqf_host_logic_replays_reference_traces(gne, name, tmp_path):
    """Through the arrays entry point, and - for the fixtures that carry their text (shuffled lines, duplicate lines, members
    of capacity 0, an empty set) - through the text loader."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    n, p = int(z["n"]), int(z["p"])
    col = None
    if "col_text" in z.files:
        col = str(tmp_path / "inst.col")
        open(col, "wb").write(z["col_text"].tobytes())
    checked = 0
    for rn in rules_of(z):
        r1, r2 = rn.split("+") if "+" in rn else (rn, rn)
        if r1 not in DRAWS or r2 not in DRAWS:
            continue                                    # CL2 / CLW2: the number of draws per pricing depends on the pricing itself
        p1, p2 = (int(v) for v in z[rn + "_p"])
        runs = []
        if arrays_ok(z):
            runs.append(gne.host_replay_eqf(n, z["tail"], z["head"], z["ub"], z["cost"], z["mult"], z["eqf"], p, z["supply"], -1.0,
                                            (DRAWS[r1], DRAWS[r2]), p1, z[rn + "_e"], z[rn + "_l"], len(z[rn + "_flows"])))
        if col:
            runs.append(gne.host_replay_eqf_col(col, n, (DRAWS[r1], DRAWS[r2]), p1, z[rn + "_e"], z[rn + "_l"], len(z[rn + "_flows"])))
        assert runs
        for (mm, flows, pots, obj) in runs:
            assert mm == -1, (name, rn, "leaving variable differs at pivot", mm, "of", p1, p2)
            assert same(flows, z[rn + "_flows"]), (name, rn)
            assert same(pots, z[rn + "_potentials"]), (name, rn)
            assert same_scalar(obj, float(z[rn + "_objective"])), (name, rn)
        checked += 1
    assert checked >= 2


def test_eqf_fixtures_exercise_type_two_pivots():
    """The fixtures must reach what they are for: sets entering and leaving the basis (pivotTII, gnePivot.cpp:122-160)."""
    total_in = total_out = 0
    for name in FIXTURES:
        z = np.load(os.path.join(GOLD, name + ".npz"))
        if not arrays_ok(z):
            continue
        nvars = int(np.sum((z["ub"] > 0.0) & (z["eqf"] == 0))) + int(z["n"])
        for rn in rules_of(z):
            total_in += int(np.sum(z[rn + "_e"] >= nvars))
            total_out += int(np.sum(z[rn + "_l"] >= nvars))
    assert total_in > 500 and total_out > 500


def test_eqf_path_sparse_and_full_sweeps_agree():
    """GNE_EQF_DENSE=1 turns the path-sparse Type I sweeps off (the reference's full sweeps run instead): the same replays must
    pass - the two differ only in which exact zeros they compute."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, genneteq_b200 as gne\n"
            "import glob, os\n"
            "D = dict(LV=1, LVE=1, QS=0, CL=0, FV=0, RLV=1)\n"
            "n = 0\n"
            "for f in sorted(glob.glob(os.path.join(%r, 'q_*.npz'))):\n"
            "    z = np.load(f)\n"
            "    if 'arrays_ok' in z.files and int(z['arrays_ok']) == 0: continue\n"
            "    for rn in D:\n"
            "        if rn + '_e' not in z.files: continue\n"
            "        mm, fl, po, ob = gne.host_replay_eqf(int(z['n']), z['tail'], z['head'], z['ub'], z['cost'], z['mult'], z['eqf'], int(z['p']),\n"
            "                                             z['supply'], -1.0, D[rn], int(z[rn + '_p'][0]), z[rn + '_e'], z[rn + '_l'], len(z[rn + '_flows']))\n"
            "        assert mm == -1 and np.array_equal(fl, z[rn + '_flows'], equal_nan=True) and np.array_equal(po, z[rn + '_potentials'], equal_nan=True), (f, rn)\n"
            "        n += 1\n"
            "print('DENSE-OK', n)\n" % (ROOT, GOLD))
    env = dict(os.environ)
    env["GNE_EQF_DENSE"] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DENSE-OK" in r.stdout, r.stdout[-500:] + r.stderr[-2000:]
    assert int(r.stdout.split("DENSE-OK")[1].split()[0]) >= 20


def test_eqf_no_cpu_fallback(gne):
    """Without a device an instance with sets is refused like any other (the host-only replay is test infrastructure)."""
    if gne.cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    z = np.load(os.path.join(GOLD, "q_wide_s21.npz"))
    with pytest.raises(gne.GneError) as ei:
        gne.Solver(int(z["n"]), z["tail"], z["head"], z["ub"], z["cost"], z["mult"], z["supply"], eqf=z["eqf"], num_sets=int(z["p"]))
    assert "no CPU fallback" in str(ei.value)


def test_eqf_replay_detects_a_wrong_trace(gne):
    z = np.load(os.path.join(GOLD, "q_wide_s21.npz"))
    l = z["LV_l"].copy()
    k = 40
    l[k] = l[k] + 1 if l[k] + 1 != z["LV_e"][k] else l[k] + 2
    mm, *_ = gne.host_replay_eqf(int(z["n"]), z["tail"], z["head"], z["ub"], z["cost"], z["mult"], z["eqf"], int(z["p"]), z["supply"],
                                 -1.0, 1, int(z["LV_p"][0]), z["LV_e"], l, len(z["LV_flows"]))
    assert mm == k


def test_eqf_arguments_are_validated(gne):
    z = np.load(os.path.join(GOLD, "q_wide_s21.npz"))
    bad = z["eqf"].copy()
    bad[3] = int(z["p"]) + 1                            # names a set beyond the problem line's count
    with pytest.raises(gne.GneError):
        gne.host_replay_eqf(int(z["n"]), z["tail"], z["head"], z["ub"], z["cost"], z["mult"], bad, int(z["p"]), z["supply"],
                            -1.0, 1, 0, z["LV_e"][:0], z["LV_l"][:0], len(z["LV_flows"]))


def test_eqf_text_loader_edge_cases(gne, tmp_path):
    """q_edge_s31: what the loader makes of a duplicate line, a member of capacity 0, an empty set (gneInit.cpp:64-103)."""
    z = np.load(os.path.join(GOLD, "q_edge_s31.npz"))
    col = tmp_path / "inst.col"
    col.write_bytes(z["col_text"].tobytes())
    r = gne.read_instance_eqf(str(col))
    assert r["num_sets"] == 6 and r["n"] == int(z["n"])
    lines = int(z["tail"].shape[0])
    assert len(r["tail"]) == lines - 2                  # the duplicate collapses, the member of capacity 0 is no arc
    assert not np.any(r["eqf"] == 5)                    # nobody names set 5
    assert np.sum(r["eqf"] == 4) == 2                   # three lines name set 4, one of them with capacity 0
    assert len(z["LV_flows"]) == int(np.sum(r["eqf"] == 0)) + r["n"] + 6


@pytest.mark.parametrize("name", ["q_wide_s21", "q_near1_s24"])
def test_eqf_text_reader_returns_the_set_column(gne, tmp_path, name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    col = tmp_path / "inst.col"
    col.write_bytes(z["col_text"].tobytes())
    r = gne.read_instance_eqf(str(col))
    assert (r["n"], r["num_sets"]) == (int(z["n"]), int(z["p"]))
    keep = z["ub"] > 0.0
    for k in ("tail", "head", "ub", "cost", "mult", "eqf"):
        assert np.array_equal(r[k], z[k][keep]), k
    assert np.array_equal(r["supply"], z["supply"])
    with pytest.raises(gne.GneError) as ei:
        gne.read_instance(str(col))
    assert ei.value.code == gne.GNE_ERR_UNSUPPORTED


REFERENCE_WENT_ASTRAY = {("q_lossy_s22", "LVE")}       # sets-first Dantzig on this instance: 7 484 Phase II pivots to an objective of -2.8e6


def test_lu_restatement_against_lapack(gne):
    """The s x s solve (GSL's LU order restated): same pivot rows as LAPACK's partial pivoting (first row of largest
    magnitude), P A = L U to rounding, solution within 1e-10 of numpy's - on random, badly scaled and tie-heavy systems."""
    import scipy.linalg as sl
    rng = np.random.default_rng(3)
    for s in (1, 2, 3, 7, 20, 64):
        for kind in ("normal", "scaled", "ties"):
            a = rng.normal(size=(s, s))
            if kind == "scaled":
                a *= 10.0 ** rng.integers(-6, 6, size=(s, 1))
            if kind == "ties":
                a = rng.choice([-2.0, -1.0, 0.5, 1.0, 2.0], size=(s, s)) + 3.0 * np.eye(s)
            b = rng.normal(size=s)
            x, lu, perm = gne.host_lu_solve(a, b)
            assert np.allclose(x, np.linalg.solve(a, b), rtol=1e-9, atol=1e-10 * np.abs(x).max())
            L = np.tril(lu, -1) + np.eye(s); U = np.triu(lu)
            assert np.allclose(L @ U, a[perm], rtol=1e-12, atol=1e-12 * np.abs(a).max())
            assert np.all(np.abs(np.tril(lu, -1)) <= 1.0 + 1e-15)                # partial pivoting: multipliers <= 1
            if kind != "ties":                          # (with exact ties LAPACK's blocked kernels may pick another of the tied rows)
                _, piv = sl.lu_factor(a)
                q = np.arange(s)
                for i, r in enumerate(piv):
                    q[[i, r]] = q[[r, i]]
                assert np.array_equal(q, perm)


@pytest.mark.parametrize("name", [f for f in FIXTURES if f != "q_edge_s31"])
def test_eqf_optimum_against_an_independent_lp_solver(name):
    """SURVEY section 4 (iv), the dense cross-check: the instance as a plain LP - one column per arc outside the sets, ONE column
    per set (sum of its member arcs' node-arc columns, capacity = the smallest member capacity, cost = the members' sum) - solved
    by scipy's HiGHS.  The reference's optimum (through the restated GSL LU) equals it, and the fixture's final flows satisfy the
    conservation rows and bounds: the equal-flow semantics are pinned independently of any LU."""
    import scipy.optimize as so
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLD, name + ".npz"))
    n, p = int(z["n"]), int(z["p"])
    tail, head, ub, cost, mult, eqf, sup = (z[k] for k in ("tail", "head", "ub", "cost", "mult", "eqf", "supply"))
    keep = ub > 0
    free = np.flatnonzero(keep & (eqf == 0))
    nv = len(free) + p
    rows, cols, vals = [], [], []
    c = np.zeros(nv); U = np.zeros(nv)
    for j, a in enumerate(free):
        rows += [tail[a], head[a]]; cols += [j, j]; vals += [1.0, -mult[a]]
        c[j] = cost[a]; U[j] = ub[a]
    for r in range(p):
        j = len(free) + r
        mem = np.flatnonzero(keep & (eqf == r + 1))
        for a in mem:
            rows += [tail[a], head[a]]; cols += [j, j]; vals += [1.0, -mult[a]]
            c[j] += cost[a]
        U[j] = ub[mem].min()
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, nv))
    res = so.linprog(c, A_eq=A, b_eq=sup, bounds=np.column_stack([np.zeros(nv), U]), method="highs")
    verdicts = [int(z[rn + "_feasible"]) for rn in rules_of(z)]
    if not any(verdicts):
        # every run of the reference ended Phase I infeasible (no Phase II pivots): the LP solver must say so too
        assert res.status == 2, (name, res.status)
        assert all(int(z[rn + "_p"][1]) == 0 for rn in rules_of(z))
        return
    assert res.status == 0
    checked = 0
    for rn in rules_of(z):
        obj = float(z[rn + "_objective"])
        if not np.isfinite(obj) or not int(z[rn + "_feasible"]):
            continue                                    # runs in which the reference itself ended in NaN / infeasible after Phase I
        x = z[rn + "_flows"]
        xa = np.concatenate([x[:len(free)], x[len(free) + n:]])                # arc variables, (artificial self-loops), sets
        if (name, rn) in REFERENCE_WENT_ASTRAY:
            # the reference's own run left the feasible region (its flows break the bounds): kept as a parity fixture only
            assert xa.min() < -1.0 or (xa - U).max() > 1.0
            continue
        assert abs(obj - res.fun) <= 1e-7 * max(1.0, abs(res.fun)), (name, rn, obj, res.fun)
        assert np.abs(A @ xa - sup).max() < 1e-7 and np.abs(x[len(free):len(free) + n]).max() < 1e-9
        assert xa.min() > -1e-9 and (xa - U).max() < 1e-9
        checked += 1
    assert checked >= 2, name


# ------------------------------------------------------------------------------------------------------------------
# GPU: the solve itself
# ------------------------------------------------------------------------------------------------------------------
def solve_fixture(gne, z, rn, tmp_path=None, **kw):
    r1, r2 = rn.split("+") if "+" in rn else (rn, rn)
    if not arrays_ok(z):                                # only the text describes the instance (duplicate lines)
        col = str(tmp_path / "inst.col")
        open(col, "wb").write(z["col_text"].tobytes())
        return gne.Solver(col_path=col, rule1=r1, rule2=r2, **kw)
    s = gne.Solver(int(z["n"]), z["tail"], z["head"], z["ub"], z["cost"], z["mult"], z["supply"], rule1=r1, rule2=r2,
                   eqf=z["eqf"], num_sets=int(z["p"]), **kw)
    return s


def check_against(z, rn, s, p1, p2):
    assert (p1, p2) == tuple(int(v) for v in z[rn + "_p"]), rn
    e, l = s.trace
    assert np.array_equal(e, z[rn + "_e"]) and np.array_equal(l, z[rn + "_l"]), rn
    assert same_scalar(s.objective, float(z[rn + "_objective"])), rn
    assert same(s.flows(), z[rn + "_flows"]), rn
    assert same(s.potentials(), z[rn + "_potentials"]), rn
    assert s.feasible == bool(int(z[rn + "_feasible"])), rn


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_eqf_solve_equals_reference_every_rule(gne, name, tmp_path):
    """Arcs priced by the LV / QS / first-violator / compaction kernels, sets by k_price_eqf, selection state carried from
    one to the other on the host: identical pivots, flows, potentials and objective for every rule of the fixture."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    for rn in rules_of(z):
        s = solve_fixture(gne, z, rn, tmp_path)
        assert s.num_sets == int(z["p"])
        p1, p2 = s.solve_both()
        check_against(z, rn, s, p1, p2)
        c = s.counters()
        assert c["launches"] > 0 and c["arcs_priced"] > 0
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["q_wide_s21", "q_near1_s24"])
def test_eqf_text_loader(gne, tmp_path, name):
    """The reference's text format with sets (graph.cpp:382-419): d_r(i) summed in FILE order (the second fixture's `a` lines
    are shuffled), variables numbered in (tail, head) order."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    col = tmp_path / "inst.col"
    col.write_bytes(z["col_text"].tobytes())
    for rn in ("LV", "QS"):
        s = gne.Solver(col_path=str(col), rule1=rn)
        assert s.num_sets == int(z["p"]) and s.n == int(z["n"])
        p1, p2 = s.solve_both()
        check_against(z, rn, s, p1, p2)
        s.close()


@pytest.mark.gpu
def test_eqf_unsupported_pieces_say_so(gne):
    z = np.load(os.path.join(GOLD, "q_wide_s21.npz"))
    s = solve_fixture(gne, z, "SD")
    with pytest.raises(gne.GneError) as ei:
        s.solve(True)
    assert ei.value.code == gne.GNE_ERR_UNSUPPORTED
    s.close()
    s = solve_fixture(gne, z, "LV")
    with pytest.raises(gne.GneError):
        s.forest()
    s.close()


@pytest.mark.gpu
def test_eqf_pivot_budget_and_resume(gne):
    """A solve cut by the pivot budget continues to the same trace."""
    z = np.load(os.path.join(GOLD, "q_pow2_s23.npz"))
    s = solve_fixture(gne, z, "LV")
    it, fin = s.solve(True, 50)
    assert (it, fin) == (50, False)
    e, l = s.trace
    assert np.array_equal(e, z["LV_e"][:50]) and np.array_equal(l, z["LV_l"][:50])
    s.close()
