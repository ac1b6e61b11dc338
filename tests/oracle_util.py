"""ctypes bindings of the checkers under oracle/ (TEST INFRASTRUCTURE).

``Oracle``  - the CPU restatement oracle/liboracle.so (any number of instances per process).
``RefProc`` - the unmodified reference, oracle/_ref/libgne_ref.so; Tree keeps static pointers
              (trees.h:83-86), so each use runs in a fresh subprocess.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgne_ref.so")
REF_DRIVER = os.path.join(ORACLE_DIR, "_ref", "gne_ref_driver")
REF_MAIN = os.path.join(ORACLE_DIR, "_ref", "genneteq")

RULES = dict(LV=0, LVW=1, LVA=2, LVE=3, FV=4, SD=5, RV=6, RWV=7, RLV=8, RLVW=9, CL=10, CLW=11,
             SAA=12, SAE=13, CL2=14, CLW2=15, QS=16, QSW=17)

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_long)
c_bp = C.POINTER(C.c_byte)


def dp(a):
    return a.ctypes.data_as(c_dp)


def ip(a):
    return a.ctypes.data_as(c_ip)


def lp(a):
    return a.ctypes.data_as(c_lp)


def bp(a):
    return a.ctypes.data_as(c_bp)


GEN_SO = os.path.join(ORACLE_DIR, "libgne_gen.so")
REF_WINDOWS = os.path.join(ROOT, "tests", "golden", "ref_windows.npz")
GEN_MODES = dict(wide=0, lossy=1, near1=2, pow2=3)


def generate_native(n, m, seed, mode="lossy"):
    """The product's synthetic-instance generator through the stand-alone oracle/libgne_gen.so (same source file,
    same flags => bit-identical instances) - for processes that must not map the product library."""
    if not os.path.exists(GEN_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "gen"], stdout=subprocess.DEVNULL)
    gen = C.CDLL(GEN_SO)
    tail = np.empty(m, np.int32); head = np.empty(m, np.int32)
    ub = np.empty(m); cost = np.empty(m); mult = np.empty(m); supply = np.empty(n)
    rc = gen.gne_generate_instance_impl(C.c_int(n), C.c_int64(m), C.c_uint64(seed), C.c_int(GEN_MODES[mode]),
                                        tail.ctypes.data_as(C.POINTER(C.c_int32)), head.ctypes.data_as(C.POINTER(C.c_int32)),
                                        dp(ub), dp(cost), dp(mult), dp(supply))
    if rc != 0:
        raise ValueError("gne_generate_instance_impl(%d, %d): status %d" % (n, m, rc))
    return dict(n=n, tail=tail, head=head, ub=ub, cost=cost, mult=mult, supply=supply)


def ref_window(name):
    """(e, l, md5) of a prefix window written by the unmodified reference (tests/golden/make_ref_windows.py)."""
    z = np.load(REF_WINDOWS)
    return z[name + "_e"], z[name + "_l"], str(z[name + "_md5"])


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "gne_oracle.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.orc_create.restype = C.c_void_p
        lib.orc_big_m.restype = C.c_double
        lib.orc_solve.restype = C.c_long
        lib.orc_trace_len.restype = C.c_long
        lib.orc_objective.restype = C.c_double
        lib.orc_num_vars.restype = C.c_long
        lib.orc_num_nonbasic.restype = C.c_long
        lib.orc_drand48.restype = C.c_double
        lib.orc_drand48_initial_state.restype = C.c_ulonglong
        lib.orc_price_lv.restype = C.c_long
        lib.orc_price_violators.restype = C.c_long
        lib.orc_flow_apply.restype = C.c_double
        _lib = lib
    return _lib


class Instance:
    """Arcs sorted by (tail, head); float64 data; 0-based int32 node ids."""

    def __init__(self, n, tail, head, ub, cost, mult, supply):
        self.n = int(n)
        self.tail = np.ascontiguousarray(tail, np.int32)
        self.head = np.ascontiguousarray(head, np.int32)
        self.ub = np.ascontiguousarray(ub, np.float64)
        self.cost = np.ascontiguousarray(cost, np.float64)
        self.mult = np.ascontiguousarray(mult, np.float64)
        self.supply = np.ascontiguousarray(supply, np.float64)
        self.m = len(self.tail)

    @property
    def big_m(self):
        return oracle_lib().orc_big_m(C.c_long(self.m), dp(self.cost), dp(self.ub))

    @staticmethod
    def from_col(path):
        from gen_instance import read_col
        d = read_col(path)
        return Instance(d["n"], d["tail"], d["head"], d["ub"], d["cost"], d["mult"], d["supply"])

    @staticmethod
    def generate(n, m, seed, mode):
        """Appendix-A generator, with the values rounded through the 6-decimal text form."""
        from gen_instance import generate
        arcs, sup = generate(n, m, seed, mode)
        r6 = lambda x: float("%.6f" % x)
        return Instance(n, [a[0] for a in arcs], [a[1] for a in arcs], [r6(a[2]) for a in arcs],
                        [r6(a[3]) for a in arcs], [r6(a[4]) for a in arcs], [r6(s) for s in sup])

    def save(self, path):
        np.savez_compressed(path, n=self.n, tail=self.tail, head=self.head, ub=self.ub, cost=self.cost,
                            mult=self.mult, supply=self.supply)

    @staticmethod
    def load(path):
        z = np.load(path)
        return Instance(int(z["n"]), z["tail"], z["head"], z["ub"], z["cost"], z["mult"], z["supply"])

    def write_col(self, path):
        with open(path, "w") as f:
            f.write("p efpgn %d %d 0\n" % (self.n, self.m))
            for a in range(self.m):
                f.write("a %d %d 0.0 %.6f %.6f %.6f 0\n" % (self.tail[a] + 1, self.head[a] + 1, self.ub[a],
                                                          self.cost[a], self.mult[a]))
            for i in range(self.n):
                f.write("n %d %.6f\n" % (i + 1, self.supply[i]))


def trace_text(e, l, p1):
    out = []
    it = 0
    for i in range(len(e)):
        if i == p1:
            it = 0
        out.append("P %d %d %d\n" % (it, e[i], l[i]))
        it += 1
    return "".join(out)


def trace_md5(e, l, p1):
    return hashlib.md5(trace_text(e, l, p1).encode()).hexdigest()


class Oracle:
    def __init__(self, inst, rule1=0, rule2=None, trace_cap=4_000_000):
        lib = oracle_lib()
        self.lib = lib
        self.inst = inst
        self.h = C.c_void_p(lib.orc_create(inst.n, C.c_long(inst.m), ip(inst.tail), ip(inst.head), dp(inst.ub),
                                           dp(inst.cost), dp(inst.mult), dp(inst.supply), C.c_double(inst.big_m)))
        lib.orc_set_rules(self.h, rule1, rule1 if rule2 is None else rule2)
        self.te = np.zeros(trace_cap, np.int32)
        self.tl = np.zeros(trace_cap, np.int32)
        lib.orc_set_trace(self.h, ip(self.te), ip(self.tl), C.c_long(trace_cap))
        self.M = inst.m + inst.n

    def solve(self, presolve, max_pivots=-1):
        fin = C.c_int(0)
        it = self.lib.orc_solve(self.h, 1 if presolve else 0, C.c_long(max_pivots), C.byref(fin))
        return it, bool(fin.value)

    def solve_both(self):
        p1, f1 = self.solve(True)
        p2, f2 = self.solve(False)
        return p1, p2

    @property
    def trace(self):
        k = self.lib.orc_trace_len(self.h)
        return self.te[:k].copy(), self.tl[:k].copy()

    @property
    def status(self):
        return self.lib.orc_status(self.h)

    @property
    def objective(self):
        return self.lib.orc_objective(self.h)

    @property
    def feasible(self):
        return bool(self.lib.orc_feasible(self.h))

    def flows(self):
        a = np.zeros(self.M)
        self.lib.orc_get_flows(self.h, dp(a))
        return a

    def potentials(self):
        a = np.zeros(self.inst.n)
        self.lib.orc_get_potentials(self.h, dp(a))
        return a

    def compute_potentials(self):
        self.lib.orc_compute_potentials(self.h)
        return self.potentials()

    def states(self):
        a = np.zeros(self.M, np.int32)
        self.lib.orc_get_states(self.h, ip(a))
        return a

    def costs(self):
        a = np.zeros(self.M)
        self.lib.orc_get_costs(self.h, dp(a))
        return a

    def nonbasic_order(self):
        a = np.zeros(self.lib.orc_num_nonbasic(self.h), np.int32)
        self.lib.orc_get_nonbasic_order(self.h, ip(a))
        return a

    def forest(self):
        n = self.inst.n
        arrs = [np.zeros(n, np.int32) for _ in range(5)]
        self.lib.orc_get_forest(self.h, *[ip(a) for a in arrs])
        return dict(zip(("tree", "pred", "predArc", "depth", "thread"), arrs))

    def timers(self):
        a = np.zeros(5)
        self.lib.orc_get_timers(self.h, dp(a))
        return dict(pricing=a[0], potentials=a[1], pivot=a[2], total=a[3], arcs_priced=a[4])

    def nonbasic_soa(self):
        """The nonbasic structure-of-arrays in setNBarc position order (what the device mirrors)."""
        inst = self.inst
        allt = np.concatenate([inst.tail, np.arange(inst.n, dtype=np.int32)])
        allh = np.concatenate([inst.head, np.arange(inst.n, dtype=np.int32)])
        allm = np.concatenate([inst.mult, np.where(inst.supply >= 0, 0.5, 2.0)])
        order = self.nonbasic_order()
        st = self.states()
        cst = self.costs()
        return dict(order=order, tail=allt[order].copy(), head=allh[order].copy(), cost=cst[order].copy(),
                    mult=allm[order].copy(), state=st[order].astype(np.int8))

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_REF_CHILD = r"""
import ctypes as C, json, sys, numpy as np
sys.path.insert(0, %(tests)r)
from oracle_util import Instance, dp, ip, REF_SO
req = json.load(open(sys.argv[1]))
inst = Instance.load(req["instance"])
ref = C.CDLL(REF_SO)
ref.ref_solve.restype = C.c_long; ref.ref_trace_len.restype = C.c_long; ref.ref_objective.restype = C.c_double
ref.ref_num_vars.restype = C.c_long
ref.ref_feasible.restype = C.c_int
ref.ref_init_sparse(inst.n, C.c_long(inst.m), ip(inst.tail), ip(inst.head), dp(inst.ub), dp(inst.cost), dp(inst.mult),
                    dp(inst.supply), C.c_double(req["big_m"]))
ref.ref_set_rules(req["rule1"], req["rule2"])
cap = req["trace_cap"]
te = np.zeros(cap, np.int32); tl = np.zeros(cap, np.int32)
ref.ref_set_trace(ip(te), ip(tl), C.c_long(cap))
fin = C.c_int(0)
p1 = ref.ref_solve(1, C.c_long(req["max_pivots"]), C.byref(fin)); f1 = fin.value
p2 = 0; f2 = 0
if f1:
    p2 = ref.ref_solve(0, C.c_long(req["max_pivots"]), C.byref(fin)); f2 = fin.value
k = ref.ref_trace_len()
M = inst.m + inst.n
x = np.zeros(M); pi = np.zeros(inst.n); st = np.zeros(M, np.int32)
ref.ref_get_flows(dp(x)); ref.ref_get_potentials(dp(pi)); ref.ref_get_states(ip(st))
tm = np.zeros(5); ref.ref_get_timers(dp(tm))
np.savez(req["out"], e=te[:k], l=tl[:k], p1=p1, p2=p2, f1=f1, f2=f2, objective=ref.ref_objective(), flows=x,
         potentials=pi, states=st, timers=tm, feasible=ref.ref_feasible())
"""


def run_reference(inst, rule1, rule2=None, max_pivots=-1, trace_cap=4_000_000, tmpdir="/tmp"):
    """Solve with the unmodified reference (sparse-initialised) in a subprocess; returns an npz dict."""
    import tempfile
    d = tempfile.mkdtemp(prefix="gne_ref_", dir=tmpdir)
    ipath = os.path.join(d, "inst.npz")
    inst.save(ipath)
    req = dict(instance=ipath, big_m=inst.big_m, rule1=rule1, rule2=rule1 if rule2 is None else rule2,
               max_pivots=max_pivots, trace_cap=trace_cap, out=os.path.join(d, "out.npz"))
    rpath = os.path.join(d, "req.json")
    json.dump(req, open(rpath, "w"))
    code = _REF_CHILD % dict(tests=os.path.join(ROOT, "tests"))
    subprocess.check_call([sys.executable, "-c", code, rpath], stdout=subprocess.DEVNULL)
    z = dict(np.load(req["out"]))
    import shutil
    shutil.rmtree(d, ignore_errors=True)
    return z


def have_reference():
    return os.path.exists(REF_SO) and os.path.exists(REF_DRIVER)
