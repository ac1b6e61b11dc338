"""Drop-in boundary, executed: the UNMODIFIED reference host engine (compiled from /root/reference into
oracle/_ref/gne_gpu_driver by oracle/Makefile) with only its pricing pass routed through the product's C ABI
(layer 1: gne_nb_reset / gne_nb_write / gne_pot_set_all / gne_price_lv / gne_price_qs), as INTEGRATION.md
section A describes.  The pivot trace must equal the reference's own golden trace.

Reference: GenNetEq::solve genneteq.cpp:60-117, getVarWithLargestViolation gnePivotRules.cpp:133-157,
quickSelect gnePivotRules.cpp:710-764, updateBasis gnePivot.cpp:254-309.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

from oracle_util import RULES  # noqa: E402
import gen_instance  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "gne_gpu_driver")
GOLDEN = os.path.join(HERE, "golden")


def _run_driver(col, r1, r2, trace_path, extra=()):
    out = subprocess.run([DRIVER, "-1", str(r1), "-2", str(r2), "-T", trace_path, *extra, col],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    piv = []
    with open(trace_path) as f:
        for line in f:
            if line.startswith("P "):
                _, it, e, l = line.split()
                piv.append((int(it), int(e), int(l)))
    obj = None
    for line in out.stdout.splitlines():
        if line.startswith("GPU-REF pivots"):
            obj = float(line.split()[-1])
    _run_driver.last_stdout = out.stdout
    return np.array(piv, dtype=np.int64).reshape(-1, 3), obj


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,rule", [
    ("s_wide_s3", "LV"), ("s_wide_s3", "QS"), ("s_lossy_s4", "LV"), ("s_pow2_s5", "LV"),
    ("s_near1_s6", "QSW"), ("c1_wide_s7", "LV"), ("c1_wide_s7", "QS"),
    ("e_infeasible_s11", "LV"), ("e_nosupply_s12", "QS"), ("e_two_nodes_s13", "LV"),
])
def test_reference_engine_with_gpu_pricing(fixture, rule, tmp_path):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/gne_gpu_driver not built (needs /root/reference at build time)")
    z = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    if rule + "_e" not in z.files:
        pytest.skip("no golden trace for %s/%s" % (fixture, rule))
    n, m, seed = (int(v) for v in z["params"])
    arcs, sup = gen_instance.apply_tweak(*gen_instance.generate(n, m, seed, str(z["mode"])), str(z["tweak"]) if "tweak" in z.files else "")
    col = str(tmp_path / (fixture + ".col"))
    gen_instance.write_col(col, n, arcs, sup)
    assert hashlib.md5(open(col, "rb").read()).hexdigest() == str(z["col_md5"])
    r = RULES[rule]
    trace = str(tmp_path / "trace.txt")
    piv, obj = _run_driver(col, r, r, trace)
    assert np.array_equal(piv[:, 1], z[rule + "_e"]) and np.array_equal(piv[:, 2], z[rule + "_l"])
    assert hashlib.md5(open(trace, "rb").read()).hexdigest() == str(z[rule + "_md5"])
    assert obj == float(z[rule + "_objective"])


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,rule,extra", [("q_wide_s21", "LV", ()), ("q_wide_s21", "LVW", ()), ("q_wide_s21", "QS", ()), ("q_near1_s24", "QSW", ()),
                                                ("q_near1_s24", "LV", ("-A", "3")), ("q_edge_s31", "QS", ("-A", "2"))])
def test_reference_engine_with_gpu_pricing_and_equal_flow_sets(fixture, rule, extra, tmp_path):
    """Mode A on instances WITH equal-flow sets: the reference's own engine (Type II trees, pivotTII, its s x s solves) with the
    arcs priced by gne_price_lv / gne_price_qs and the sets by gne_price_eqf reproduces the reference's own trace; with -A the
    reference's selectEnteringVariable() and the GPU path are compared from the same state as well."""
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/gne_gpu_driver not built (needs /root/reference at build time)")
    z = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    col = str(tmp_path / (fixture + ".col"))
    open(col, "wb").write(z["col_text"].tobytes())
    r = RULES[rule]
    trace = str(tmp_path / "trace.txt")
    piv, obj = _run_driver(col, r, r, trace, extra)
    assert np.array_equal(piv[:, 1], z[rule + "_e"]) and np.array_equal(piv[:, 2], z[rule + "_l"])
    assert hashlib.md5(open(trace, "rb").read()).hexdigest() == str(z[rule + "_md5"])
    assert obj == float(z[rule + "_objective"])
    if extra:
        ab = [ln for ln in _run_driver.last_stdout.splitlines() if ln.startswith("A/B")]
        assert ab, _run_driver.last_stdout[-500:]
        f = ab[-1].split()
        assert int(f[2]) > 10 and int(f[4]) == 0, ab[-1]            # "A/B samples N mismatches M ..."


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,rule,every", [("c1_wide_s7", "LV", 5), ("c1_wide_s7", "QS", 3), ("s_lossy_s4", "LV", 1),
                                                ("e_nosupply_s12", "QS", 1)])
def test_same_state_ab_against_the_reference_rule(fixture, rule, every, tmp_path):
    """SURVEY 8(d)(ii), literally: inside the reference's own solve loop, at every `every`-th iteration the
    reference's selectEnteringVariable() (host) and the GPU path price the SAME state (drand48 state, cursor and flags
    saved / restored in between); the two entering variables must be the same object every time, and the extra calls
    must not disturb the pivot sequence (golden md5)."""
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/gne_gpu_driver not built (needs /root/reference at build time)")
    z = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    n, m, seed = (int(v) for v in z["params"])
    arcs, sup = gen_instance.apply_tweak(*gen_instance.generate(n, m, seed, str(z["mode"])), str(z["tweak"]) if "tweak" in z.files else "")
    col = str(tmp_path / (fixture + ".col"))
    gen_instance.write_col(col, n, arcs, sup)
    trace = str(tmp_path / "trace.txt")
    r = RULES[rule]
    piv, obj = _run_driver(col, r, r, trace, extra=("-A", str(every)))
    line = [ln for ln in _run_driver.last_stdout.splitlines() if ln.startswith("A/B samples")][-1].split()
    samples, mismatches = int(line[2]), int(line[4])
    print(" ".join(line))
    assert samples >= len(piv) // every and mismatches == 0
    assert hashlib.md5(open(trace, "rb").read()).hexdigest() == str(z[rule + "_md5"])
    assert obj == float(z[rule + "_objective"])
