"""Host-side logic of the product, checked on CPU (no GPU needed): the C ABI surface, the basis
forest engine against the oracle's forest, the shard-record merge, the instance generator."""
import ctypes as C
import os
import sys
import re
import subprocess

import numpy as np
import pytest

from oracle_util import RULES, Instance, Oracle, oracle_lib, dp, ip, lp, bp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gne():
    import __graft_entry__ as ge
    ge.build()
    import genneteq_b200
    return genneteq_b200


def test_library_exports_every_declared_symbol(gne):
    hdr = open(os.path.join(ROOT, "include", "genneteq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gne_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gne_status"}
    assert len(declared) >= 45
    nm = subprocess.check_output(["nm", "-D", "--defined-only", gne.LIB_PATH], text=True)
    exported = set(re.findall(r" T (gne_[a-z0-9_]+)", nm))
    assert declared <= exported, sorted(declared - exported)
    assert declared == set(gne.SIGNATURES), sorted(declared ^ set(gne.SIGNATURES))
    assert b"sm_100a" in gne.lib.gne_version()


def test_no_cpu_fallback_without_a_device(gne):
    if gne.cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(gne.GneError) as ei:
        gne.Device(10, 100)
    assert "no CPU fallback" in str(ei.value)
    inst = Instance.load(os.path.join(GOLD, "c1_wide_s7_instance.npz"))
    with pytest.raises(gne.GneError):
        gne.Solver(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply)


def test_product_does_not_link_or_name_the_oracle(gne):
    src = os.path.join(ROOT, "genneteq_b200")
    for dirpath, _, files in os.walk(src):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "gne_oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("the oracle's", ""), f
    ldd = subprocess.check_output(["ldd", gne.LIB_PATH], text=True)
    assert "oracle" not in ldd


@pytest.mark.parametrize("name,rule,K", [("c1_wide_s7", "LV", None), ("c1_wide_s7", "QS", 700), ("s_pow2_s5", "CL", None),
                                         ("s_lossy_s4", "FV", None), ("lossy3k_s5", "LV", 6000)])
def test_forest_engine_matches_oracle_forest(gne, name, rule, K):
    """Replay the reference's pivot trace through the product's forest engine: the thread, pred,
    predArc and depth arrays must equal the oracle's after the same exchanges."""
    from test_oracle_pinned import load
    z, inst = load(name)
    e = z[rule + "_e"]; l = z[rule + "_l"]
    p1 = int(z[rule + "_p"][0])
    K = p1 if K is None else min(K, p1)
    o = Oracle(inst, RULES[rule])
    o.solve(True, K)
    fo = o.forest()
    fp = gne.forest_replay(inst.n, inst.tail, inst.head, inst.supply, e[:K], l[:K])
    for key in ("pred", "predArc", "depth", "thread"):
        assert np.array_equal(fo[key], fp[key]), key
    # tree ids differ (stable slots vs renumbered ids) but must induce the same partition
    pairs = set(zip(fo["tree"].tolist(), fp["tree"].tolist()))
    assert len(pairs) == len(set(fo["tree"].tolist())) == len(set(fp["tree"].tolist()))
    o.close()


def test_merge_lv_shards_equals_sequential_rule(gne):
    """Per-shard (max, in-band candidates) records merged by gne_merge_lv_shards give the same
    largestViolation and offending list as the sequential band rule over the whole list."""
    lib = oracle_lib()
    inst = Instance.load(os.path.join(GOLD, "c1_wide_s7_instance.npz"))
    band = 4.0e-13
    for K in (0, 5, 300, 1500):
        o = Oracle(inst, RULES["LV"])
        o.solve(True, K)
        pi = o.compute_potentials()
        nb = o.nonbasic_soa()
        N = len(nb["order"])
        lst = np.zeros(N, np.int64); L = C.c_double(0); anyv = C.c_int(0)
        k = lib.orc_price_lv(C.c_long(N), ip(nb["tail"]), ip(nb["head"]), dp(nb["cost"]), dp(nb["mult"]), bp(nb["state"]),
                             dp(pi), lp(lst), C.c_long(N), C.byref(L), C.byref(anyv))
        allpos = np.zeros(N, np.int64); allv = np.zeros(N)
        cnt = lib.orc_price_violators(C.c_long(N), ip(nb["tail"]), ip(nb["head"]), dp(nb["cost"]), dp(nb["mult"]),
                                      bp(nb["state"]), dp(pi), C.c_double(0.0), lp(allpos), dp(allv), C.c_long(N))
        allpos = allpos[:cnt]; allv = allv[:cnt]
        for R in (1, 2, 3, 8):
            per = ((N + R - 1) // R + 2047) // 2048 * 2048
            stride = 64
            maxima = np.full(R, -1.0); anys = np.zeros(R, np.int32); counts = np.zeros(R, np.int64)
            lp_ = np.zeros(R * stride, np.int64); lv_ = np.zeros(R * stride)
            for r in range(R):
                sel = (allpos >= r * per) & (allpos < (r + 1) * per)
                if sel.any():
                    anys[r] = 1
                    maxima[r] = allv[sel].max()
                    keep = sel & (allv >= maxima[r] - band)
                    c = int(keep.sum())
                    counts[r] = c if c <= stride else -1
                    lp_[r * stride: r * stride + min(c, stride)] = allpos[keep][:stride]
                    lv_[r * stride: r * stride + min(c, stride)] = allv[keep][:stride]
            res = gne.merge_lv_shards(maxima, anys, counts, lp_, lv_, stride)
            assert res["any"] == bool(anyv.value)
            if not res["need_fallback"]:
                assert res["largest"] == L.value
                assert np.array_equal(res["list"], lst[:k])
        o.close()


def test_generator_is_deterministic_sorted_and_connected(gne):
    a = gne.generate_instance(500, 6000, 99, "lossy")
    b = gne.generate_instance(500, 6000, 99, "lossy")
    for k in a:
        assert np.array_equal(a[k], b[k])
    key = a["tail"].astype(np.int64) * 500 + a["head"]
    assert np.all(np.diff(key) > 0)                      # sorted by (tail, head), no duplicates
    assert np.all(a["tail"] != a["head"])
    ring = set(zip(range(500), [(i + 1) % 500 for i in range(500)]))
    assert ring <= set(zip(a["tail"].tolist(), a["head"].tolist()))
    assert a["mult"].min() >= 0.70 and a["mult"].max() <= 0.99
    assert np.all(np.round(a["cost"] * 1e6) / 1e6 == a["cost"])
    c = gne.generate_instance(500, 6000, 100, "lossy")
    assert not np.array_equal(a["head"], c["head"])


DRAWS = dict(LV=1, LVW=1, LVA=1, LVE=1, FV=0, RV=1, RWV=1, RLV=1, RLVW=1, CL=0, CLW=0, SAA=2, SAE=2, QS=0, QSW=0)


@pytest.mark.parametrize("sparse", [True, False])
@pytest.mark.parametrize("name,rules", [("c1_wide_s7", ["LV", "QS", "CL"]),
                                        ("s_wide_s3", ["LV", "FV", "RV", "RWV", "RLV", "CL", "SAA", "QS"]),
                                        ("s_lossy_s4", ["LV", "FV", "RWV", "CL", "SAE", "QS"]),
                                        ("s_pow2_s5", ["LV", "FV", "RV", "RLV", "CL", "QS"]),
                                        ("s_near1_s6", ["LV", "FV", "RWV", "CL", "SAA", "QS"]),
                                        ("e_infeasible_s11", ["LV", "FV", "CL", "QS", "RV"]),
                                        ("e_nosupply_s12", ["LV", "FV", "CL", "QS", "RV"]),
                                        ("e_two_nodes_s13", ["LV", "QS"]),
                                        ("e_three_nodes_s14", ["LV", "FV", "RV"])])
def test_host_pivot_logic_replays_reference_traces(gne, name, rules, sparse):
    """Flows, ratio test, leaving arc, basis update and the potential sweep of the product's host engine,
    driven by the reference's entering arcs (no pricing, no GPU): every leaving arc, the final flows, the
    final potentials and the objective must equal the reference's."""
    from test_oracle_pinned import load
    z, inst = load(name)
    for rn in rules:
        e = z[rn + "_e"]; l = z[rn + "_l"]
        p1 = int(z[rn + "_p"][0])
        mm, flows, pots, obj = gne.host_replay(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply,
                                               inst.big_m, DRAWS[rn], p1, e, l, sparse_flows=sparse)
        assert mm == -1, (rn, "leaving arc differs at pivot", mm)
        assert obj == float(z[rn + "_objective"]), rn
        assert np.array_equal(flows, z[rn + "_flows"]), rn
        assert np.array_equal(pots, z[rn + "_potentials"]), rn


def test_host_replay_lossy3k(gne):
    from test_oracle_pinned import load
    z, inst = load("lossy3k_s5")
    mm, flows, pots, obj = gne.host_replay(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply,
                                           inst.big_m, 1, int(z["LV_p"][0]), z["LV_e"], z["LV_l"])
    assert mm == -1
    assert obj == float(z["LV_objective"])


def test_instance_files_round_trip(gne, tmp_path):
    """Text format as the reference reads / writes it (graph.cpp:94-121, 358-447) and the binary SoA form."""
    from test_oracle_pinned import load
    z, inst = load("s_wide_s3")
    d = dict(n=inst.n, tail=inst.tail, head=inst.head, ub=inst.ub, cost=inst.cost, mult=inst.mult, supply=inst.supply)
    col = str(tmp_path / "i.col"); binp = str(tmp_path / "i.bin")
    gne.write_instance(col, d)
    gne.write_instance(binp, d, binary=True)
    a = gne.read_instance(col); b = gne.read_instance(binp, binary=True)
    for k in ("tail", "head", "ub", "cost", "mult", "supply"):
        assert np.array_equal(a[k], d[k]), k
        assert np.array_equal(b[k], d[k]), k
    assert a["n"] == b["n"] == inst.n
    assert a["big_m"] == inst.big_m                       # graph.cpp:43
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gne_ref_driver")):   # the reference's own loader accepts the file
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "gne_ref_driver"), "-1", "16", "-2", "16", col],
                           stdout=subprocess.PIPE, text=True)
        last = [ln for ln in r.stdout.splitlines() if ln.startswith("REF pivots")][-1].split()
        assert (int(last[2]), int(last[3])) == tuple(int(v) for v in z["QS_p"])
    # quirks of the reference's parser: unsorted input is sorted, a later duplicate wins, capacity <= 0 drops the arc
    q = str(tmp_path / "q.col")
    open(q, "w").write("c x\np efpgn 3 4 0\na 2 3 0.0 5.0 1.0 1.0 0\na 1 2 0.0 5.0 2.0 0.9 0\na 1 2 0.0 7.0 3.0 0.8 0\na 3 1 0.0 0.0 9.0 1.0 0\nn 1 1\nn 2 0\nn 3 -1\n")
    g = gne.read_instance(q)
    assert g["tail"].tolist() == [0, 1] and g["head"].tolist() == [1, 2]
    assert g["ub"].tolist() == [7.0, 5.0] and g["cost"].tolist() == [3.0, 1.0]
    assert g["big_m"] == 4 * 9.0 * 7.0


def test_periodic_flow_recompute_skips_only_unchanged_trees():
    """computeFlows (every 1000 iterations, gneFlow.cpp:30-56) walks only the nonbasic arcs at their upper bound and
    re-sweeps only the trees a pivot, a structure change or a U-arc move touched.  With GNE_FLOW_CHECK=1 the engine
    runs the reference's full walk next to it at every call and aborts on the first differing bit; the 15 464-pivot
    3k-node trace (15 periodic recomputes per phase boundary) and configs[0] must replay cleanly."""
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import genneteq_b200 as gne
from test_oracle_pinned import load
for name, rule in (("lossy3k_s5", "LV"), ("c1_wide_s7", "QS"), ("c1_wide_s7", "LV"), ("e_infeasible_s11", "FV")):
    z, inst = load(name)
    draws = dict(LV=1, QS=0, FV=0)[rule]
    mm, flows, pots, obj = gne.host_replay(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply,
                                           inst.big_m, draws, int(z[rule + "_p"][0]), z[rule + "_e"], z[rule + "_l"])
    assert mm == -1 and obj == float(z[rule + "_objective"]), (name, rule, mm)
print("FLOW-CHECK-OK")
""" % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"))
    env = dict(os.environ)
    env["GNE_FLOW_CHECK"] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FLOW-CHECK-OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


@pytest.mark.parametrize("text,code", [
    ("p efpgn x 2 0\n", 2),                                                     # improperly formatted problem line
    ("a 1 2 0.0 5.0 1.0 1.0 0\n", 2),                                           # arc before the problem line
    ("p efpgn 2 1 0\na 1 3 0.0 5.0 1.0 1.0 0\nn 1 1\nn 2 -1\n", 2),             # arc names a node out of range
    ("p efpgn 2 2 0\na 1 2 0.0 5.0 1.0 1.0 0\nn 1 1\nn 2 -1\n", 2),             # fewer arcs than announced
    ("p efpgn 2 1 0\na 1 2 0.0 5.0 1.0\nn 1 1\nn 2 -1\n", 2),                   # short arc line
    ("p efpgn 2 1 0\na 1 2 0.0 5.0 1.0 1.0 0\nn 7 1\n", 2),                     # node line out of range
    ("p efpgn 2 1 1\na 1 2 0.0 5.0 1.0 1.0 1\nn 1 1\nn 2 -1\n", 6),             # equal-flow sets: gne_read_col has no column for them
    ("p efpgn 2 1 1\na 1 2 0.0 5.0 1.0 1.0 2\nn 1 1\nn 2 -1\n", 2),             # arc names a set beyond the problem line's count
    ("p efpgn 2 1 -3\n", 2),                                                   # negative set count
    ("p efpgn -1 0 0\n", 2),                                                    # negative node count (no exception across the C ABI)
    ("p efpgn 0 0 0\n", 2),
    ("p efpgn 3 -5 0\n", 2),                                                    # negative arc count
    ("p efpgn 2000000000 1 0\n", 2),                                            # absurd node count
])
def test_col_parser_rejects_malformed_files(gne, tmp_path, text, code):
    """Where the reference prints and exits (graph.cpp:358-447) the loader returns a status code and a message."""
    p = str(tmp_path / "bad.col")
    open(p, "w").write(text)
    with pytest.raises(gne.GneError) as ei:
        gne.read_instance(p)
    assert ei.value.code == code, (ei.value.code, str(ei.value))
    with pytest.raises(gne.GneError):
        gne.read_instance(str(tmp_path / "missing.col"))


def test_bin_header_is_validated_against_the_file_length(gne, tmp_path):
    """gne_read_bin_header returns sizes callers allocate from: a header that does not match the file is refused."""
    import struct
    p = str(tmp_path / "bad.bin")
    open(p, "wb").write(b"GNESOA1\0" + struct.pack("<qq", -4, 10))
    with pytest.raises(gne.GneError):
        gne.read_instance(p, binary=True)
    open(p, "wb").write(b"GNESOA1\0" + struct.pack("<qq", 1 << 29, 1 << 40))
    with pytest.raises(gne.GneError):
        gne.read_instance(p, binary=True)


@pytest.mark.parametrize("m,R", [(100, 2), (10_000, 8), (2047, 8), (2048, 2), (2049, 2), (60_000, 4), (20_000_000, 8), (1, 8), (0, 3)])
def test_shard_ranges_partition_the_list(gne, m, R):
    """Ranges are tile-aligned, disjoint, in rank order and cover [0, m + slack); ranks without positions own an
    empty shard (overlapping ranges duplicated candidates when m < 2048 x ranks)."""
    rng = [gne.shard_range(m, r, R) for r in range(R)]
    covered = 0
    for lo, hi in rng:
        assert lo % 2048 == 0 and hi >= lo
        if hi > lo:
            assert lo == covered, rng
            covered = hi
    assert covered == m + 16, rng


def test_merge_rejects_overlapping_shards(gne):
    """Duplicate positions from overlapping shards must not silently inflate the tie list."""
    import numpy as np
    stride = 4
    maxima = np.array([5.0, 5.0]); anys = np.array([1, 1], np.int32); counts = np.array([1, 1], np.int64)
    pos = np.zeros(2 * stride, np.int64); viol = np.zeros(2 * stride)
    pos[0] = 7; viol[0] = 5.0; pos[stride] = 7; viol[stride] = 5.0
    with pytest.raises(gne.GneError):
        gne.merge_lv_shards(maxima, anys, counts, pos, viol, stride)
    pos[stride] = 9
    r = gne.merge_lv_shards(maxima, anys, counts, pos, viol, stride)
    assert list(r["list"]) == [7, 9] and r["largest"] == 5.0


def test_reference_arm_generator_library_matches_the_product(gne):
    """bench.py's reference arm builds its workload from oracle/libgne_gen.so (so that the reference process
    never maps the product library); both are compiled from the same source with the same flags."""
    import ctypes as C
    import numpy as np
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "gen"], stdout=subprocess.DEVNULL)
    g = C.CDLL(os.path.join(ROOT, "oracle", "libgne_gen.so"))
    for (n, m, seed, mode) in [(1000, 10_000, 7, "wide"), (5000, 100_000, 21, "lossy"), (300, 4000, 3, "pow2")]:
        a = gne.generate_instance(n, m, seed, mode)
        t = np.empty(m, np.int32); h = np.empty(m, np.int32); ub = np.empty(m); c = np.empty(m); mu = np.empty(m); s = np.empty(n)
        ip = C.POINTER(C.c_int32); dp = C.POINTER(C.c_double)
        rc = g.gne_generate_instance_impl(n, C.c_int64(m), C.c_uint64(seed), dict(wide=0, lossy=1, near1=2, pow2=3)[mode],
                                          t.ctypes.data_as(ip), h.ctypes.data_as(ip), ub.ctypes.data_as(dp), c.ctypes.data_as(dp),
                                          mu.ctypes.data_as(dp), s.ctypes.data_as(dp))
        assert rc == 0
        for k, v in (("tail", t), ("head", h), ("ub", ub), ("cost", c), ("mult", mu), ("supply", s)):
            assert np.array_equal(a[k], v), k


def test_host_replay_deep_c2_trace(gne):
    """Deep into a configs[1]-size solve (100k nodes / 2M arcs, 56 000 Dantzig pivots: one basis tree of ~50k nodes, cut
    subtrees of thousands of nodes): the host engine - subtree-cut split with the implicit arcsInTree order, partial
    re-threading, sparse flows - must pick every recorded leaving arc.  The recorded trace itself was checked against the
    UNMODIFIED reference (tests/golden/verify_deep_fixture.py: its own pivot() reproduces every leaving arc, and its own
    solve continues with exactly the recorded pivots)."""
    from oracle_util import generate_native, Instance
    z = np.load(os.path.join(ROOT, "tests", "golden", "deep_c2_lv.npz"))
    g = generate_native(int(z["n"]), int(z["m"]), int(z["seed"]), str(z["mode"]))
    inst = Instance(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"])
    e = np.ascontiguousarray(z["e"]); l = np.ascontiguousarray(z["l"])
    mm, flows, pots, obj = gne.host_replay(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply, inst.big_m,
                                           1, len(e), e, l)
    assert mm == -1, ("leaving arc differs at pivot", mm)
