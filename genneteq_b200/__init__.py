"""genneteq_b200 — B200 (sm_100a) pricing / relabel path of GenNetEq's generalized network simplex.

This module is a thin ctypes binding of ``genneteq_b200/lib/libgenneteq_b200.so`` (C ABI in
``include/genneteq_b200.h``).  There is no CPU fallback: importing fails loudly when the library
has not been built (``python -c "import __graft_entry__ as g; g.build()"``), and every compute
call fails with ``GneError`` when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgenneteq_b200.so")

RULES = dict(LV=0, LVW=1, LVA=2, LVE=3, FV=4, SD=5, RV=6, RWV=7, RLV=8, RLVW=9, CL=10, CLW=11,
             SAA=12, SAE=13, CL2=14, CLW2=15, QS=16, QSW=17)          # main.h:17-34
AT_LOWER, AT_UPPER = 1, 2
TOL = 1.0e-13

if os.environ.get("GNE_LIB"):                            # a differently built library (make EXTRA=-DGNE_STAMPS OUT=...: tools/stamps_probe.py)
    LIB_PATH = os.path.join(_HERE, "lib", os.environ["GNE_LIB"])
if os.environ.get("GNE_TUNING_LIB") == "1":              # tools/variants.py: the library built with `make tuning`
    LIB_PATH = os.path.join(_HERE, "lib", "libgenneteq_b200_tuning.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(
        "genneteq_b200: %s is missing - build it with `make -C genneteq_b200/csrc` "
        "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/genneteq_b200.h
SIGNATURES = {
    "gne_last_error": (C.c_char_p, []),
    "gne_version": (C.c_char_p, []),
    "gne_cuda_device_count": (C.c_int, []),
    "gne_dev_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64]),
    "gne_dev_destroy": (C.c_int, [_vp]),
    "gne_dev_sync": (C.c_int, [_vp]),
    "gne_nb_reset": (C.c_int, [_vp, C.c_int64, _i32p, _i32p, _f64p, _f64p, _i8p]),
    "gne_nb_set_costs": (C.c_int, [_vp, C.c_int64, _f64p]),
    "gne_nb_write": (C.c_int, [_vp, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int8]),
    "gne_nb_set_count": (C.c_int, [_vp, C.c_int64]),
    "gne_nb_read": (C.c_int, [_vp, C.c_int64, C.c_int64, _i32p, _i32p, _f64p, _f64p, _i8p]),
    "gne_pot_update_nodes": (C.c_int, [_vp, C.c_int64, _i32p, _f64p, _f64p, _i32p]),
    "gne_pot_update_thetas": (C.c_int, [_vp, C.c_int64, _i32p, _f64p]),
    "gne_pot_finalize": (C.c_int, [_vp]),
    "gne_pot_finalize_slots": (C.c_int, [_vp, C.c_int, _i32p]),
    "gne_pot_update_fused": (C.c_int, [_vp, C.c_int64, _i32p, _f64p, _f64p, _i32p, C.c_int64, _i32p, _f64p, C.c_int64, _i32p]),
    "gne_pot_update_fused_pi": (C.c_int, [_vp, C.c_int64, _i32p, _f64p, _f64p, _i32p, C.c_int64, _i32p, _f64p, C.c_int64, _i32p, _f64p]),
    "gne_pot_set": (C.c_int, [_vp, C.c_int64, _i32p, _f64p]),
    "gne_pot_set_all": (C.c_int, [_vp, _f64p]),
    "gne_pot_get_all": (C.c_int, [_vp, _f64p]),
    "gne_price_lv": (C.c_int, [_vp, _f64p, _i64p, _i64p, C.c_int64, C.POINTER(C.c_int)]),
    "gne_price_lv_begin": (C.c_int, [_vp, _f64p, _i64p, C.POINTER(C.c_int)]),
    "gne_price_lv_launch": (C.c_int, [_vp]),
    "gne_price_lv_ready": (C.c_int, [_vp]),
    "gne_price_lv_complete": (C.c_int, [_vp, _f64p, _i64p, C.POINTER(C.c_int)]),
    "gne_price_lv_at": (C.c_int, [_vp, C.c_int64, _i64p]),
    "gne_price_qs": (C.c_int, [_vp, C.c_int64, C.c_int64, _i64p, _f64p, _i64p, _i64p, _i64p]),
    "gne_price_scan_list": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, _i64p, _i64p, _f64p, C.c_int64, _i64p, _i64p]),
    "gne_price_violators": (C.c_int, [_vp, C.c_double, _i64p, _i64p, _f64p, C.c_int64]),
    "gne_price_first": (C.c_int, [_vp, _i64p]),
    "gne_price_eqf": (C.c_int, [_vp, C.c_int, _i64p, _i32p, _f64p, _f64p, _f64p]),
    "gne_reduced_costs": (C.c_int, [_vp, C.c_int64, C.c_int64, _f64p]),
    "gne_dev_last_kernel_ms": (C.c_int, [_vp, _f32p]),
    "gne_dev_set_timing": (C.c_int, [_vp, C.c_int]),
    "gne_dev_set_lv_kernel": (C.c_int, [_vp, C.c_int]),
    "gne_dev_last_lv_kernel": (C.c_int, [_vp]),
    "gne_dev_set_fast_math": (C.c_int, [_vp, C.c_int]),
    "gne_dev_set_stream_ctas": (C.c_int, [_vp, C.c_int]),
    "gne_dev_launch_count": (C.c_int64, [_vp]),
    "gne_bench_price_lv": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32p, _f32p]),
    "gne_bench_price_lv_burst": (C.c_int, [_vp, C.c_int, C.c_int, _f32p]),
    "gne_bench_price_qs": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _f32p, _i64p, _i64p]),
    "gne_bench_lv_groups": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i64p]),
    "gne_lv_group_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int]),
    "gne_lv_group_add": (C.c_int, [_vp, _vp]),
    "gne_lv_group_launch": (C.c_int, [_vp]),
    "gne_lv_group_pending": (C.c_int, [_vp]),
    "gne_lv_group_launch_count": (C.c_int64, [_vp]),
    "gne_lv_group_destroy": (C.c_int, [_vp]),
    "gne_dev_cuda_device": (C.c_int, [_vp]),
    "gne_solver_solve_batch_grouped": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, _i64p, C.POINTER(C.c_int), _i64p]),
    "gne_comm_unique_id": (C.c_int, [_u8p]),
    "gne_comm_init": (C.c_int, [_vp, _u8p, C.c_int, C.c_int]),
    "gne_comm_init_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gne_comm_destroy": (C.c_int, [_vp]),
    "gne_comm_mode": (C.c_int, [_vp]),
    "gne_solver_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, C.c_double]),
    "gne_solver_create_eqf": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, _i32p, C.c_int,
                                        _f64p, C.c_double]),
    "gne_solver_num_sets": (C.c_int, [_vp]),
    "gne_eqf_set_rows": (C.c_int, [_vp, C.c_int, _i64p, _i32p, _f64p]),
    "gne_price_eqf_launch": (C.c_int, [_vp, C.c_int, _i32p, _f64p, C.c_int64]),
    "gne_price_eqf_fetch": (C.c_int, [_vp, C.c_int, _f64p]),
    "gne_host_replay_eqf_col": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, _i32p, _i32p, _i64p, C.c_int64, _f64p, _f64p,
                                          _f64p]),
    "gne_host_lu_solve": (C.c_int, [C.c_int, _f64p, _f64p, _f64p, _f64p, _i32p]),
    "gne_host_replay_eqf": (C.c_int, [C.c_int, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, _i32p, C.c_int, _f64p, C.c_double, C.c_int, C.c_int,
                                      C.c_int64, C.c_int64, _i32p, _i32p, _i64p, _f64p, _f64p, _f64p]),
    "gne_solver_create_from_col": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_char_p]),
    "gne_solver_destroy": (C.c_int, [_vp]),
    "gne_solver_set_rules": (C.c_int, [_vp, C.c_int, C.c_int]),
    "gne_solver_set_comm": (C.c_int, [_vp, _u8p, C.c_int, C.c_int]),
    "gne_solver_set_comm_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gne_solver_set_trace": (C.c_int, [_vp, _i32p, _i32p, C.c_int64]),
    "gne_solver_trace_len": (C.c_int64, [_vp]),
    "gne_solver_solve_batch": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int64, C.c_int, _i64p, C.POINTER(C.c_int)]),
    "gne_solver_solve": (C.c_int, [_vp, C.c_int, C.c_int64, _i64p, C.POINTER(C.c_int)]),
    "gne_solver_solve_batch_ex": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, _i64p,
                                            C.POINTER(C.c_int), _i64p]),
    "gne_solver_num_nodes": (C.c_int, [_vp]),
    "gne_solver_num_vars": (C.c_int64, [_vp]),
    "gne_solver_objective": (C.c_double, [_vp]),
    "gne_solver_feasible": (C.c_int, [_vp]),
    "gne_solver_get_flows": (C.c_int, [_vp, _f64p]),
    "gne_solver_get_potentials": (C.c_int, [_vp, _f64p]),
    "gne_solver_compute_potentials": (C.c_int, [_vp]),
    "gne_solver_get_states": (C.c_int, [_vp, _i32p]),
    "gne_solver_get_forest": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "gne_solver_get_counters": (C.c_int, [_vp, _f64p]),
    "gne_solver_device": (_vp, [_vp]),
    "gne_solver_set_window": (C.c_int, [_vp, C.c_int64, C.c_int64]),
    "gne_solver_get_window": (C.c_int, [_vp, _f64p]),
    "gne_solver_set_progress": (C.c_int, [_vp, C.c_int64, _f64p, C.c_int64]),
    "gne_solver_progress_rows": (C.c_int64, [_vp]),
    "gne_forest_replay": (C.c_int, [C.c_int, C.c_int64, _i32p, _i32p, _f64p, C.c_int64, _i32p, _i32p,
                                    _i32p, _i32p, _i32p, _i32p, _i32p]),
    "gne_host_replay": (C.c_int, [C.c_int, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, C.c_double, C.c_int, C.c_int,
                                  C.c_int64, C.c_int64, _i32p, _i32p, _i64p, _f64p, _f64p, _f64p]),
    "gne_shard_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, _i64p, _i64p]),
    "gne_merge_lv_shards": (C.c_int, [C.c_int, _f64p, C.POINTER(C.c_int), _i64p, _i64p, _f64p, C.c_int64,
                                      _f64p, _i64p, _i64p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gne_write_col": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]),
    "gne_read_col": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), _i64p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p]),
    "gne_read_col_eqf": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), _i64p, C.POINTER(C.c_int), _i32p, _i32p, _f64p, _f64p, _f64p, _i32p,
                                   _f64p, _f64p]),
    "gne_write_bin": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]),
    "gne_read_bin_header": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), _i64p]),
    "gne_read_bin": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]),
    "gne_generate_instance": (C.c_int, [C.c_int, C.c_int64, C.c_uint64, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


(GNE_OK, GNE_ERR_CUDA, GNE_ERR_ARG, GNE_ERR_CAPACITY, GNE_ERR_TREE, GNE_ERR_CYCLING, GNE_ERR_UNSUPPORTED, GNE_ERR_COMM) = range(8)


class GneError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib.gne_last_error()
        RuntimeError.__init__(self, "%s failed with status %d: %s" % (where, code, msg.decode() if msg else ""))


def _ck(code, where):
    if code != 0:
        raise GneError(code, where)


def _p(a, t):
    return a.ctypes.data_as(t)


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def cuda_device_count():
    return lib.gne_cuda_device_count()


def generate_instance(n, m, seed, mode="lossy"):
    """Deterministic large synthetic instance (arcs sorted by (tail, head)); see gne_generate_instance."""
    mode_id = dict(wide=0, lossy=1, near1=2, pow2=3)[mode]
    tail = np.empty(m, np.int32); head = np.empty(m, np.int32)
    ub = np.empty(m); cost = np.empty(m); mult = np.empty(m); supply = np.empty(n)
    _ck(lib.gne_generate_instance(n, m, seed, mode_id, _p(tail, _i32p), _p(head, _i32p), _p(ub, _f64p), _p(cost, _f64p),
                                  _p(mult, _f64p), _p(supply, _f64p)), "gne_generate_instance")
    return dict(n=n, tail=tail, head=head, ub=ub, cost=cost, mult=mult, supply=supply)


def write_instance(path, inst, binary=False):
    """inst: dict(n, tail, head, ub, cost, mult, supply); text (reference format) or binary SoA."""
    t = _arr(inst["tail"], np.int32); h = _arr(inst["head"], np.int32); ub = _arr(inst["ub"], np.float64)
    c = _arr(inst["cost"], np.float64); mu = _arr(inst["mult"], np.float64); s = _arr(inst["supply"], np.float64)
    fn = lib.gne_write_bin if binary else lib.gne_write_col
    _ck(fn(path.encode(), int(inst["n"]), len(t), _p(t, _i32p), _p(h, _i32p), _p(ub, _f64p), _p(c, _f64p), _p(mu, _f64p),
           _p(s, _f64p)), "gne_write_bin" if binary else "gne_write_col")


def read_instance(path, binary=False):
    n = C.c_int(0); m = C.c_int64(0); big = C.c_double(0)
    if binary:
        _ck(lib.gne_read_bin_header(path.encode(), C.byref(n), C.byref(m)), "gne_read_bin_header")
    else:
        _ck(lib.gne_read_col(path.encode(), C.byref(n), C.byref(m), None, None, None, None, None, None, C.byref(big)), "gne_read_col")
    t = np.empty(m.value, np.int32); h = np.empty(m.value, np.int32)
    ub = np.empty(m.value); c = np.empty(m.value); mu = np.empty(m.value); s = np.empty(n.value)
    if binary:
        _ck(lib.gne_read_bin(path.encode(), n.value, m.value, _p(t, _i32p), _p(h, _i32p), _p(ub, _f64p), _p(c, _f64p),
                             _p(mu, _f64p), _p(s, _f64p)), "gne_read_bin")
    else:
        _ck(lib.gne_read_col(path.encode(), C.byref(n), C.byref(m), _p(t, _i32p), _p(h, _i32p), _p(ub, _f64p), _p(c, _f64p),
                             _p(mu, _f64p), _p(s, _f64p), C.byref(big)), "gne_read_col")
    out = dict(n=n.value, tail=t, head=h, ub=ub, cost=c, mult=mu, supply=s)
    if not binary:
        out["big_m"] = big.value
    return out


def read_instance_eqf(path):
    """A text instance that may hold equal-flow sets (gne_read_col_eqf): read_instance's dict plus eqf, num_sets."""
    n = C.c_int(0); m = C.c_int64(0); p = C.c_int(0); big = C.c_double(0)
    _ck(lib.gne_read_col_eqf(path.encode(), C.byref(n), C.byref(m), C.byref(p), None, None, None, None, None, None, None,
                             C.byref(big)), "gne_read_col_eqf")
    t = np.empty(m.value, np.int32); h = np.empty(m.value, np.int32); q = np.empty(m.value, np.int32)
    ub = np.empty(m.value); c = np.empty(m.value); mu = np.empty(m.value); s = np.empty(n.value)
    _ck(lib.gne_read_col_eqf(path.encode(), C.byref(n), C.byref(m), C.byref(p), _p(t, _i32p), _p(h, _i32p), _p(ub, _f64p),
                             _p(c, _f64p), _p(mu, _f64p), _p(q, _i32p), _p(s, _f64p), C.byref(big)), "gne_read_col_eqf")
    return dict(n=n.value, tail=t, head=h, ub=ub, cost=c, mult=mu, supply=s, eqf=q, num_sets=p.value, big_m=big.value)


class Device:
    """Layer 1: the device operators on a nonbasic structure-of-arrays in position order."""

    def __init__(self, n, capacity, device=0, shard=None):
        self.h = _vp()
        lo, hi = shard if shard is not None else (0, capacity)
        _ck(lib.gne_dev_create(C.byref(self.h), device, n, capacity, lo, hi), "gne_dev_create")
        self.n = n
        self.capacity = capacity
        self.N = 0

    def close(self):
        if self.h:
            lib.gne_dev_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nb_reset(self, tail, head, cost, mult, state):
        tail = _arr(tail, np.int32); head = _arr(head, np.int32); cost = _arr(cost, np.float64)
        mult = _arr(mult, np.float64); state = _arr(state, np.int8)
        self.N = len(tail)
        _ck(lib.gne_nb_reset(self.h, self.N, _p(tail, _i32p), _p(head, _i32p), _p(cost, _f64p), _p(mult, _f64p),
                             _p(state, _i8p)), "gne_nb_reset")

    def nb_set_costs(self, cost):
        cost = _arr(cost, np.float64)
        _ck(lib.gne_nb_set_costs(self.h, len(cost), _p(cost, _f64p)), "gne_nb_set_costs")

    def nb_write(self, pos, tail, head, cost, mult, state):
        _ck(lib.gne_nb_write(self.h, pos, tail, head, cost, mult, state), "gne_nb_write")

    def nb_set_count(self, N):
        self.N = N
        _ck(lib.gne_nb_set_count(self.h, N), "gne_nb_set_count")

    def nb_read(self, lo, hi):
        c = hi - lo
        tail = np.empty(c, np.int32); head = np.empty(c, np.int32); cost = np.empty(c); mult = np.empty(c)
        state = np.empty(c, np.int8)
        _ck(lib.gne_nb_read(self.h, lo, hi, _p(tail, _i32p), _p(head, _i32p), _p(cost, _f64p), _p(mult, _f64p),
                            _p(state, _i8p)), "gne_nb_read")
        return tail, head, cost, mult, state

    def pot_set_all(self, pi):
        pi = _arr(pi, np.float64)
        assert len(pi) == self.n
        _ck(lib.gne_pot_set_all(self.h, _p(pi, _f64p)), "gne_pot_set_all")

    def pot_set(self, node, val):
        node = _arr(node, np.int32); val = _arr(val, np.float64)
        _ck(lib.gne_pot_set(self.h, len(node), _p(node, _i32p), _p(val, _f64p)), "gne_pot_set")

    def pot_get_all(self):
        pi = np.empty(self.n)
        _ck(lib.gne_pot_get_all(self.h, _p(pi, _f64p)), "gne_pot_get_all")
        return pi

    def pot_update_nodes(self, node, a0, a1, slot):
        node = _arr(node, np.int32); a0 = _arr(a0, np.float64); a1 = _arr(a1, np.float64); slot = _arr(slot, np.int32)
        _ck(lib.gne_pot_update_nodes(self.h, len(node), _p(node, _i32p), _p(a0, _f64p), _p(a1, _f64p), _p(slot, _i32p)),
            "gne_pot_update_nodes")

    def pot_update_thetas(self, slot, theta):
        slot = _arr(slot, np.int32); theta = _arr(theta, np.float64)
        _ck(lib.gne_pot_update_thetas(self.h, len(slot), _p(slot, _i32p), _p(theta, _f64p)), "gne_pot_update_thetas")

    def pot_update_fused(self, node, a0, a1, slot, tslot, theta, fnode=None):
        node = _arr(node, np.int32); a0 = _arr(a0, np.float64); a1 = _arr(a1, np.float64); slot = _arr(slot, np.int32)
        tslot = _arr(tslot, np.int32); theta = _arr(theta, np.float64)
        nf = -1 if fnode is None else len(fnode)
        fn = _arr(fnode if fnode is not None else [], np.int32)
        _ck(lib.gne_pot_update_fused(self.h, len(node), _p(node, _i32p), _p(a0, _f64p), _p(a1, _f64p), _p(slot, _i32p),
                                     len(tslot), _p(tslot, _i32p), _p(theta, _f64p), nf, _p(fn, _i32p)), "gne_pot_update_fused")

    def pot_finalize(self):
        _ck(lib.gne_pot_finalize(self.h), "gne_pot_finalize")

    def price_lv(self, cap=1 << 20):
        largest = C.c_double(0); count = C.c_int64(0); anyv = C.c_int(0)
        buf = np.empty(cap, np.int64)
        _ck(lib.gne_price_lv(self.h, C.byref(largest), C.byref(count), _p(buf, _i64p), cap, C.byref(anyv)), "gne_price_lv")
        return largest.value, buf[:count.value].copy(), bool(anyv.value)

    def price_lv_begin(self):
        largest = C.c_double(0); count = C.c_int64(0); anyv = C.c_int(0)
        _ck(lib.gne_price_lv_begin(self.h, C.byref(largest), C.byref(count), C.byref(anyv)), "gne_price_lv_begin")
        return largest.value, count.value, bool(anyv.value)

    def price_lv_at(self, index):
        pos = C.c_int64(0)
        _ck(lib.gne_price_lv_at(self.h, index, C.byref(pos)), "gne_price_lv_at")
        return pos.value

    def price_qs(self, cursor, want):
        bp = C.c_int64(0); bv = C.c_double(0); nc = C.c_int64(0); sc = C.c_int64(0); vi = C.c_int64(0)
        _ck(lib.gne_price_qs(self.h, cursor, want, C.byref(bp), C.byref(bv), C.byref(nc), C.byref(sc), C.byref(vi)),
            "gne_price_qs")
        return dict(best_pos=bp.value, best_viol=bv.value, new_cursor=nc.value, scanned=sc.value, violations=vi.value)

    def price_scan_list(self, cursor, want, min_scan=5):
        cap = max(self.N, 1)
        pos = np.empty(cap, np.int64); viol = np.empty(cap)
        count = C.c_int64(0); nc = C.c_int64(0); sc = C.c_int64(0)
        _ck(lib.gne_price_scan_list(self.h, cursor, want, min_scan, C.byref(count), _p(pos, _i64p), _p(viol, _f64p), cap,
                                    C.byref(nc), C.byref(sc)), "gne_price_scan_list")
        return pos[:count.value].copy(), viol[:count.value].copy(), nc.value, sc.value

    def price_violators(self, threshold=0.0, cap=None):
        cap = cap if cap is not None else max(self.N, 1)
        pos = np.empty(cap, np.int64); viol = np.empty(cap)
        count = C.c_int64(0)
        _ck(lib.gne_price_violators(self.h, threshold, C.byref(count), _p(pos, _i64p), _p(viol, _f64p), cap),
            "gne_price_violators")
        return pos[:count.value].copy(), viol[:count.value].copy()

    def price_first(self):
        pos = C.c_int64(0)
        _ck(lib.gne_price_first(self.h, C.byref(pos)), "gne_price_first")
        return pos.value

    def reduced_costs(self, lo=0, hi=None):
        hi = self.N if hi is None else hi
        rc = np.empty(hi - lo)
        _ck(lib.gne_reduced_costs(self.h, lo, hi, _p(rc, _f64p)), "gne_reduced_costs")
        return rc

    def price_eqf(self, row_start, node, val, cost):
        """Reduced costs of equal-flow variables given as CSR rows (gne_price_eqf)."""
        rs = _arr(row_start, np.int64); nd = _arr(node, np.int32); vl = _arr(val, np.float64); cs = _arr(cost, np.float64)
        rc = np.zeros(len(cs))
        _ck(lib.gne_price_eqf(self.h, len(cs), _p(rs, _i64p), _p(nd, _i32p), _p(vl, _f64p), _p(cs, _f64p), _p(rc, _f64p)), "gne_price_eqf")
        return rc

    def eqf_set_rows(self, row_start, node, val):
        """The rows of all equal-flow variables, kept on the device (gne_eqf_set_rows)."""
        rs = _arr(row_start, np.int64); nd = _arr(node, np.int32); vl = _arr(val, np.float64)
        _ck(lib.gne_eqf_set_rows(self.h, len(rs) - 1, _p(rs, _i64p), _p(nd, _i32p), _p(vl, _f64p)), "gne_eqf_set_rows")

    def price_eqf_resident(self, ids, cost, non_finite):
        """gne_price_eqf_launch + gne_price_eqf_fetch: reduced costs of the listed resident rows."""
        ii = _arr(ids, np.int32); cs = _arr(cost, np.float64)
        _ck(lib.gne_price_eqf_launch(self.h, len(ii), _p(ii, _i32p), _p(cs, _f64p), int(non_finite)), "gne_price_eqf_launch")
        rc = np.zeros(len(ii))
        _ck(lib.gne_price_eqf_fetch(self.h, len(ii), _p(rc, _f64p)), "gne_price_eqf_fetch")
        return rc

    def pot_finalize_slots(self, slots):
        """pi = a0 + a1 * theta[slot] for every node of the given tree slots only (gne_pot_finalize_slots)."""
        sl = _arr(slots, np.int32)
        _ck(lib.gne_pot_finalize_slots(self.h, len(sl), _p(sl, _i32p)), "gne_pot_finalize_slots")

    def set_stream_ctas(self, ctas):
        """CTAs of this handle's persistent LV pass (0 = the whole device); batches share the SMs."""
        _ck(lib.gne_dev_set_stream_ctas(self.h, int(ctas)), "gne_dev_set_stream_ctas")

    def set_fast_math(self, on):
        """Fused multiply-add in the reduced cost / potential finalize (non-parity mode); per device."""
        _ck(lib.gne_dev_set_fast_math(self.h, 1 if on else 0), "gne_dev_set_fast_math")

    def set_lv_kernel(self, mode):
        _ck(lib.gne_dev_set_lv_kernel(self.h, dict(auto=0, tiles=1, stream=2)[mode]), "gne_dev_set_lv_kernel")

    def last_lv_kernel(self):
        return {1: "tiles", 2: "stream"}[lib.gne_dev_last_lv_kernel(self.h)]

    def last_kernel_ms(self):
        ms = C.c_float(0)
        lib.gne_dev_last_kernel_ms(self.h, C.byref(ms))
        return ms.value

    def launch_count(self):
        return lib.gne_dev_launch_count(self.h)

    def bench_price_lv(self, reps, flush_l2, with_finalize=2):
        """(ms per device step, ms of the tile kernel alone), CUDA events on the handle's stream."""
        ms = np.zeros(reps, np.float32); mt = np.zeros(reps, np.float32)
        _ck(lib.gne_bench_price_lv(self.h, reps, 1 if flush_l2 else 0, int(with_finalize), _p(ms, _f32p),
                                   _p(mt, _f32p)), "gne_bench_price_lv")
        return ms, mt


    def bench_price_lv_burst(self, reps, with_finalize=2):
        """ms per launch of `reps` LV passes queued back to back between one pair of events (single GPU)."""
        ms = np.zeros(1, np.float32)
        _ck(lib.gne_bench_price_lv_burst(self.h, reps, int(with_finalize), _p(ms, _f32p)), "gne_bench_price_lv_burst")
        return float(ms[0]) / reps

    def bench_price_qs(self, reps, want, flush_l2=False):
        """`reps` quickSelect passes, device-timed: (ms per pass, arcs the kernels covered, arcs the rule consumed)."""
        ms = np.zeros(reps, np.float32); al = np.zeros(reps, np.int64); sc = np.zeros(reps, np.int64)
        _ck(lib.gne_bench_price_qs(self.h, reps, 1 if flush_l2 else 0, want, _p(ms, _f32p), _p(al, _i64p), _p(sc, _i64p)), "gne_bench_price_qs")
        return ms, al, sc

    def bench_variants(self, reps=10):
        """Tuning builds only (GNE_TUNING_LIB=1 + `make -C genneteq_b200/csrc tuning`)."""
        if not hasattr(lib, "gne_bench_variants"):
            raise RuntimeError("gne_bench_variants is only in the tuning build (make tuning; GNE_TUNING_LIB=1)")
        lib.gne_bench_variants.restype = C.c_int
        lib.gne_bench_variants.argtypes = [_vp, C.c_int, _f32p, C.c_int, C.POINTER(C.c_int)]
        lib.gne_bench_variant_name.restype = C.c_char_p
        lib.gne_bench_variant_name.argtypes = [C.c_int]
        ms = np.zeros(64, np.float32); nv = C.c_int(0)
        _ck(lib.gne_bench_variants(self.h, reps, _p(ms, _f32p), 32, C.byref(nv)), "gne_bench_variants")
        return [(lib.gne_bench_variant_name(i).decode(), float(ms[2 * i]), float(ms[2 * i + 1])) for i in range(nv.value)]


class _BorrowedDevice(Device):
    def __init__(self, handle, n, capacity, N):
        self.h = handle
        self.n = n
        self.capacity = capacity
        self.N = N

    def close(self):
        self.h = _vp()


class Solver:
    """Layer 2: the host mirror of GenNetEq (initialize / solve) driving the device operators."""

    COUNTERS = ("pricing_s", "potentials_s", "pivot_s", "total_s", "arcs_priced", "kernel_ms", "h2d_bytes",
                "d2h_bytes", "launches", "pivots", "trees_s", "flows_s")

    def __init__(self, n=None, tail=None, head=None, ub=None, cost=None, mult=None, supply=None, big_m=-1.0,
                 col_path=None, device=0, rule1="LVW", rule2=None, trace_cap=4_000_000, eqf=None, num_sets=0):
        """eqf / num_sets: equal-flow sets (eqf[a] = 1-based set of arc a, 0: none) - gne_solver_create_eqf."""
        self.h = _vp()
        if col_path is not None:
            _ck(lib.gne_solver_create_from_col(C.byref(self.h), device, col_path.encode()), "gne_solver_create_from_col")
        elif num_sets > 0:
            tail = _arr(tail, np.int32); head = _arr(head, np.int32)
            ub = _arr(ub, np.float64); cost = _arr(cost, np.float64); mult = _arr(mult, np.float64)
            supply = _arr(supply, np.float64); eqf = _arr(eqf, np.int32)
            _ck(lib.gne_solver_create_eqf(C.byref(self.h), device, n, len(tail), _p(tail, _i32p), _p(head, _i32p), _p(ub, _f64p),
                                          _p(cost, _f64p), _p(mult, _f64p), _p(eqf, _i32p), num_sets, _p(supply, _f64p), big_m),
                "gne_solver_create_eqf")
        else:
            tail = _arr(tail, np.int32); head = _arr(head, np.int32)
            ub = _arr(ub, np.float64); cost = _arr(cost, np.float64); mult = _arr(mult, np.float64)
            supply = _arr(supply, np.float64)
            _ck(lib.gne_solver_create(C.byref(self.h), device, n, len(tail), _p(tail, _i32p), _p(head, _i32p), _p(ub, _f64p),
                                      _p(cost, _f64p), _p(mult, _f64p), _p(supply, _f64p), big_m), "gne_solver_create")
        self.n = lib.gne_solver_num_nodes(self.h)
        self.M = lib.gne_solver_num_vars(self.h)
        self.num_sets = lib.gne_solver_num_sets(self.h)          # > 0: the equal-flow engine runs this instance
        r1 = RULES[rule1] if isinstance(rule1, str) else rule1
        r2 = r1 if rule2 is None else (RULES[rule2] if isinstance(rule2, str) else rule2)
        lib.gne_solver_set_rules(self.h, r1, r2)
        self._te = np.zeros(trace_cap, np.int32)
        self._tl = np.zeros(trace_cap, np.int32)
        lib.gne_solver_set_trace(self.h, _p(self._te, _i32p), _p(self._tl, _i32p), trace_cap)

    def close(self):
        if self.h:
            lib.gne_solver_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_comm(self, unique_id, rank, nranks):
        uid = _arr(unique_id, np.uint8)
        _ck(lib.gne_solver_set_comm(self.h, _p(uid, _u8p), rank, nranks), "gne_solver_set_comm")

    def set_comm_host(self, rank, nranks, allgather):
        """Shard over ranks with the caller's all-gather instead of NCCL (gne_solver_set_comm_host):
        allgather(send: bytes, nbytes) -> bytes of nranks * nbytes in rank order.  Several ranks may share a device."""
        def _cb(_ctx, send, recv, nbytes):
            try:
                out = allgather(C.string_at(send, nbytes), nbytes)
                if len(out) != nbytes * nranks:
                    return 1
                C.memmove(recv, out, len(out))
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return 1
        self._gather_cb = ALLGATHER_FN(_cb)              # keep the thunk alive as long as the solver
        _ck(lib.gne_solver_set_comm_host(self.h, rank, nranks, C.cast(self._gather_cb, C.c_void_p), None),
            "gne_solver_set_comm_host")

    def solve(self, presolve, max_pivots=-1):
        it = C.c_int64(0); fin = C.c_int(0)
        _ck(lib.gne_solver_solve(self.h, 1 if presolve else 0, max_pivots, C.byref(it), C.byref(fin)), "gne_solver_solve")
        return it.value, bool(fin.value)

    def solve_both(self):
        p1, _ = self.solve(True)
        p2, _ = self.solve(False)
        return p1, p2

    @property
    def trace(self):
        k = lib.gne_solver_trace_len(self.h)
        return self._te[:k].copy(), self._tl[:k].copy()

    @property
    def objective(self):
        return lib.gne_solver_objective(self.h)

    @property
    def feasible(self):
        return bool(lib.gne_solver_feasible(self.h))

    def flows(self):
        a = np.empty(self.M)
        _ck(lib.gne_solver_get_flows(self.h, _p(a, _f64p)), "gne_solver_get_flows")
        return a

    def potentials(self):
        a = np.empty(self.n)
        _ck(lib.gne_solver_get_potentials(self.h, _p(a, _f64p)), "gne_solver_get_potentials")
        return a

    def compute_potentials(self):
        """computePotentials() for the current basis; device potentials + nonbasic mirror brought up to date."""
        _ck(lib.gne_solver_compute_potentials(self.h), "gne_solver_compute_potentials")

    def states(self):
        a = np.empty(self.M, np.int32)
        _ck(lib.gne_solver_get_states(self.h, _p(a, _i32p)), "gne_solver_get_states")
        return a

    def forest(self):
        arrs = [np.empty(self.n, np.int32) for _ in range(5)]
        _ck(lib.gne_solver_get_forest(self.h, *[_p(a, _i32p) for a in arrs]), "gne_solver_get_forest")
        return dict(zip(("tree", "pred", "predArc", "depth", "thread"), arrs))

    def counters(self):
        a = np.zeros(12)
        lib.gne_solver_get_counters(self.h, _p(a, _f64p))
        return dict(zip(self.COUNTERS, a.tolist()))

    def set_window(self, start_pivot, num_pivots):
        lib.gne_solver_set_window(self.h, start_pivot, num_pivots)

    def window(self):
        """(seconds, counters at the window's start, counters at its end) of the last windowed solve."""
        a = np.zeros(27)
        lib.gne_solver_get_window(self.h, _p(a, _f64p))
        self.window_edges = (a[25], a[26])
        return a[0], dict(zip(self.COUNTERS, a[1:13].tolist())), dict(zip(self.COUNTERS, a[13:25].tolist()))

    def set_progress(self, every, cap_rows=100000):
        self._prog = np.zeros((cap_rows, 14))
        lib.gne_solver_set_progress(self.h, every, _p(self._prog, _f64p), cap_rows)

    def progress(self):
        """rows of {wall s, counters..., largest tree size}, one per `every` pivots."""
        k = lib.gne_solver_progress_rows(self.h)
        return self._prog[:k].copy()

    def device(self):
        h = lib.gne_solver_device(self.h)
        return _BorrowedDevice(_vp(h), self.n + 16, self.M - self.n + 16, self.M - self.n) if h else None


def solve_batch(solvers, presolve, max_pivots=-1, threads=0, per_launch=0, stats=None, align_pivot=-1):
    """gne_solver_solve for several independent solvers at once: `threads` host threads (0 = one per core) interleave them;
    per_launch > 0: the LV passes of a thread's solvers leave in launch groups of up to per_launch instances (one grid);
    align_pivot >= 0: no solver starts that pivot before all have reached it (gne_solver_solve_batch_ex).
    Returns [(iterations, finished)] per solver; stats (a dict) receives group_launches."""
    B = len(solvers)
    hs = (_vp * B)(*[s.h for s in solvers])
    it = np.zeros(B, np.int64); fin = (C.c_int * B)()
    if align_pivot >= 0:
        gl = np.zeros(1, np.int64)
        _ck(lib.gne_solver_solve_batch_ex(hs, B, 1 if presolve else 0, max_pivots, threads, per_launch, align_pivot, _p(it, _i64p), fin,
                                          _p(gl, _i64p)), "gne_solver_solve_batch_ex")
        if stats is not None:
            stats["group_launches"] = int(gl[0])
    elif per_launch > 0:
        gl = np.zeros(1, np.int64)
        _ck(lib.gne_solver_solve_batch_grouped(hs, B, 1 if presolve else 0, max_pivots, threads, per_launch, _p(it, _i64p), fin,
                                               _p(gl, _i64p)), "gne_solver_solve_batch_grouped")
        if stats is not None:
            stats["group_launches"] = int(gl[0])
    else:
        _ck(lib.gne_solver_solve_batch(hs, B, 1 if presolve else 0, max_pivots, threads, _p(it, _i64p), fin), "gne_solver_solve_batch")
    return [(int(it[i]), bool(fin[i])) for i in range(B)]


class LvGroup:
    """gne_lv_group: the LV passes of several device handles of one GPU in one launch."""

    def __init__(self, device=0, max_pending=0):
        self.h = _vp()
        _ck(lib.gne_lv_group_create(C.byref(self.h), device, max_pending), "gne_lv_group_create")

    def add(self, dev):
        _ck(lib.gne_lv_group_add(self.h, dev.h), "gne_lv_group_add")

    def launch(self):
        _ck(lib.gne_lv_group_launch(self.h), "gne_lv_group_launch")

    def pending(self):
        return int(lib.gne_lv_group_pending(self.h))

    def launch_count(self):
        return int(lib.gne_lv_group_launch_count(self.h))

    def close(self):
        if self.h:
            lib.gne_lv_group_destroy(self.h)
            self.h = _vp()


def bench_lv_groups(devs, reps, per_launch, n_streams=2):
    """Device leg of a batch: (device ms of `reps` rounds of passes over all handles, launches of the batched kernel)."""
    B = len(devs)
    hs = (_vp * B)(*[d.h for d in devs])
    ms = np.zeros(1, np.float32); ln = np.zeros(1, np.int64)
    _ck(lib.gne_bench_lv_groups(hs, B, reps, per_launch, n_streams, _p(ms, _f32p), _p(ln, _i64p)), "gne_bench_lv_groups")
    return float(ms[0]), int(ln[0])


def forest_replay(n, tail, head, supply, e, l):
    """Host-only: apply the basis exchanges of a pivot trace to a fresh forest (no GPU needed)."""
    tail = _arr(tail, np.int32); head = _arr(head, np.int32); supply = _arr(supply, np.float64)
    e = _arr(e, np.int32); l = _arr(l, np.int32)
    arrs = [np.empty(n, np.int32) for _ in range(5)]
    _ck(lib.gne_forest_replay(n, len(tail), _p(tail, _i32p), _p(head, _i32p), _p(supply, _f64p), len(e), _p(e, _i32p),
                              _p(l, _i32p), *[_p(a, _i32p) for a in arrs]), "gne_forest_replay")
    return dict(zip(("tree", "pred", "predArc", "depth", "thread"), arrs))


def host_replay(n, tail, head, ub, cost, mult, supply, big_m, draws, p1, e, l, sparse_flows=True):
    """Host-only replay of a recorded solve (no GPU): returns (first mismatching pivot or -1, flows, potentials, objective)."""
    tail = _arr(tail, np.int32); head = _arr(head, np.int32); ub = _arr(ub, np.float64); cost = _arr(cost, np.float64)
    mult = _arr(mult, np.float64); supply = _arr(supply, np.float64); e = _arr(e, np.int32); l = _arr(l, np.int32)
    mm = C.c_int64(0); obj = C.c_double(0)
    flows = np.empty(len(tail) + n); pots = np.empty(n)
    _ck(lib.gne_host_replay(n, len(tail), _p(tail, _i32p), _p(head, _i32p), _p(ub, _f64p), _p(cost, _f64p), _p(mult, _f64p),
                            _p(supply, _f64p), big_m, draws, 1 if sparse_flows else 0, p1, len(e), _p(e, _i32p), _p(l, _i32p),
                            C.byref(mm), _p(flows, _f64p), _p(pots, _f64p), C.byref(obj)), "gne_host_replay")
    return mm.value, flows, pots, obj.value


def host_replay_eqf_col(path, n, draws, p1, e, l, num_vars):
    """gne_host_replay_eqf_col: the host-only replay of an equal-flow instance read from a text file."""
    d1, d2 = draws if isinstance(draws, tuple) else (draws, draws)
    e = _arr(e, np.int32); l = _arr(l, np.int32)
    mm = C.c_int64(0); obj = C.c_double(0)
    flows = np.empty(num_vars); pots = np.empty(n)
    _ck(lib.gne_host_replay_eqf_col(path.encode(), d1, d2, p1, len(e), _p(e, _i32p), _p(l, _i32p), C.byref(mm), num_vars,
                                    _p(flows, _f64p), _p(pots, _f64p), C.byref(obj)), "gne_host_replay_eqf_col")
    return mm.value, flows, pots, obj.value


def host_lu_solve(a, b):
    """The equal-flow engine's s x s solve (gne_host_lu_solve): returns (x, lu, perm)."""
    a = _arr(a, np.float64); b = _arr(b, np.float64)
    s = len(b)
    x = np.empty(s); lu = np.empty((s, s)); perm = np.empty(s, np.int32)
    _ck(lib.gne_host_lu_solve(s, _p(a, _f64p), _p(b, _f64p), _p(x, _f64p), _p(lu, _f64p), _p(perm, _i32p)), "gne_host_lu_solve")
    return x, lu, perm


def host_replay_eqf(n, tail, head, ub, cost, mult, eqf, num_sets, supply, big_m, draws, p1, e, l, num_vars):
    """The same for an instance with equal-flow sets (gne_host_replay_eqf); num_vars = arc variables + sets;
    draws = drand48 calls per pricing, one number or (Phase I, Phase II)."""
    d1, d2 = draws if isinstance(draws, tuple) else (draws, draws)
    tail = _arr(tail, np.int32); head = _arr(head, np.int32); ub = _arr(ub, np.float64); cost = _arr(cost, np.float64)
    mult = _arr(mult, np.float64); supply = _arr(supply, np.float64); e = _arr(e, np.int32); l = _arr(l, np.int32)
    eqf = _arr(eqf, np.int32)
    mm = C.c_int64(0); obj = C.c_double(0)
    flows = np.empty(num_vars); pots = np.empty(n)
    _ck(lib.gne_host_replay_eqf(n, len(tail), _p(tail, _i32p), _p(head, _i32p), _p(ub, _f64p), _p(cost, _f64p), _p(mult, _f64p),
                                _p(eqf, _i32p), num_sets, _p(supply, _f64p), big_m, d1, d2, p1, len(e), _p(e, _i32p), _p(l, _i32p),
                                C.byref(mm), _p(flows, _f64p), _p(pots, _f64p), C.byref(obj)), "gne_host_replay_eqf")
    return mm.value, flows, pots, obj.value


def shard_range(m, rank, nranks):
    """[lo, hi) of the nonbasic list that `rank` prices (gne_shard_range)."""
    lo = C.c_int64(0); hi = C.c_int64(0)
    _ck(lib.gne_shard_range(m, rank, nranks, C.byref(lo), C.byref(hi)), "gne_shard_range")
    return lo.value, hi.value


def merge_lv_shards(maxima, anys, counts, list_pos, list_viol, stride, cap=1 << 16):
    """Host-only: what every rank does with the all-gathered per-shard LV records."""
    maxima = _arr(maxima, np.float64); anys = _arr(anys, np.int32); counts = _arr(counts, np.int64)
    list_pos = _arr(list_pos, np.int64); list_viol = _arr(list_viol, np.float64)
    largest = C.c_double(0); count = C.c_int64(0); anyv = C.c_int(0); need = C.c_int(0)
    out = np.empty(cap, np.int64)
    _ck(lib.gne_merge_lv_shards(len(maxima), _p(maxima, _f64p), _p(anys, C.POINTER(C.c_int)), _p(counts, _i64p),
                                _p(list_pos, _i64p), _p(list_viol, _f64p), stride, C.byref(largest), C.byref(count),
                                _p(out, _i64p), cap, C.byref(anyv), C.byref(need)), "gne_merge_lv_shards")
    return dict(largest=largest.value, list=out[:count.value].copy(), any=bool(anyv.value), need_fallback=bool(need.value))
