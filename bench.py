#!/usr/bin/env python3
"""Benchmark of the GenNetEq pricing / relabel hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3-lv|c2-lv|c2-qs|c1-lv]

A step is one simplex iteration of the hot path: relabel the potentials the last pivot made dirty,
price every nonbasic arc, pick the entering arc, pivot.  The default workload is BASELINE.json's
configs[2] (the one its target is quoted on): 1M nodes / 20M arcs, `lossy` gains, Dantzig rule
(reference rule LV), Phase I from the initial basis; for N > 1 the nonbasic list is sharded by
contiguous position range across the N GPUs with an NCCL exchange of the per-shard record
(fixed total work: "scaling": "strong").

  value  arcs priced per second with everything resident in HBM: K device steps (the relabel kernel of
         a pivot of this window, replayed + pricing pass + reduction [+ exchange]) timed with CUDA
         events on the launching stream, max over ranks.  The pass streams 508 MB > 126 MB of L2.
  e2e    the same metric through the public C ABI (gne_solver_solve) over K real pivots after W
         warm-up pivots: host basis-forest work, the host->device copy of each step's dirty
         potentials / set updates and the device->host read of the pricing result are all inside
         the timed region.
  roofline  the pricing kernel alone (k_price_lv_stream, the TMA-streamed persistent kernel, for shards of
         >= ~2.4M arcs; k_price_lv_tiles below that): (25 B x arcs + 8 B x nodes) / its CUDA-event time
         against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the unmodified reference (oracle/_ref, built from /root/reference by oracle/Makefile)
         timed on this box's host cores on a bounded number of pivots of the same workload.

`--impl reference` times the reference's own CPU solver on the same workload (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (nodes, arcs, seed, mode, rule, description)
    "c3-lv": (1_000_000, 20_000_000, 31, "lossy", "LV", "configs[2]: 1M nodes / 20M arcs, lossy gains, seed 31, Dantzig rule (LV)"),
    "c2-lv": (100_000, 2_000_000, 21, "lossy", "LV", "configs[1] size: 100k nodes / 2M arcs, lossy gains, seed 21, Dantzig rule (LV)"),
    "c2-qs": (100_000, 2_000_000, 21, "lossy", "QS", "configs[1]: 100k nodes / 2M arcs, lossy gains, seed 21, block rule (QS)"),
    "c5-lv": (200_000, 100_000_000, 51, "lossy", "LV", "configs[4]: 200k nodes / 100M arcs, lossy gains, seed 51, Dantzig rule (LV)"),
    "c1-lv": (1_000, 10_000, 7, "wide", "LV", "configs[0] size: 1k nodes / 10k arcs (native generator), Dantzig rule (LV)"),
    # configs[3]: independent instances, one solver handle (= one CUDA stream) each, `--batch` per GPU, seeds 1000..
    "c4-batch": (50_000, 1_000_000, 1000, "lossy", "LV", "configs[3]: independent 50k-node / 1M-arc lossy instances (seeds 1000+), one stream per instance, Dantzig rule (LV)"),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the pricing kernel, read from the committed `ncu --set full`
# summary of the same workload / kernel under profiles/ (tools/ncu_summary.py writes the "traffic_bytes" line); None when
# no capture of that (workload, GPUs, kernel) is committed
def ncu_traffic(workload, world, kernel):
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path)).get("%s/n%d/%s" % (workload, world, kernel))
    except Exception:
        return None
METRIC = "arcs priced/sec"
UNIT = "arcs/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dbg(msg):
    if os.environ.get("GNE_BENCH_DEBUG"):
        sys.stderr.write("[bench %s %.1f] %s\n" % (os.environ.get("RANK", "0"), time.time() % 1000, msg))
        sys.stderr.flush()


def workload_config(wl, gpus, batch=None, window="start"):
    """The `config` object - identical in the b200 arm and the reference arm (it describes the workload, not the arm)."""
    n, m, seed, mode, rule, desc = WORKLOADS[wl]
    per_gpu = (25.0 * m / max(gpus, 1) + 8.0 * n) if batch is None else batch * (25.0 * m + 8.0 * n)
    cfg = {"workload": desc, "nodes": n, "arcs": m, "rule": rule,
           "window": ("Phase I from the initial basis: the timed steps are the pivots after the warm-up pivots" if window == "start" else
                      "deep state: Phase I re-entered at the basis reached after the recorded pivots of tests/golden/deep_%s.npz; "
                      "the timed steps are the pivots after the warm-up pivots of the re-entered solve" % wl.replace("-", "_")),
           "l2": ("inputs larger than L2 (%.0f MB streamed per pass per GPU)" % (per_gpu / 1e6) if per_gpu >= 200e6 else
                  "L2 flushed between timed device steps (320 MB write)")}
    if batch is not None:
        cfg["instances_per_gpu"] = batch
        cfg["instances"] = batch * max(gpus, 1)
    return cfg


REF_WINDOW_OF = {"c3-lv": "c3_lv", "c2-lv": "c2_lv", "c2-qs": "c2_qs", "c4-batch": "c4_lv"}


def parity_block(wl, e, l, dist, rank, world, expect=None):
    """The pivots of the e2e window (and its warm-up) checked two ways: every rank holds the same (entering, leaving)
    sequence, and the sequence equals the prefix the UNMODIFIED reference wrote for this workload
    (tests/golden/ref_windows.npz, produced by tests/golden/make_ref_windows.py from oracle/_ref)."""
    import hashlib
    e = np.ascontiguousarray(e, np.int32); l = np.ascontiguousarray(l, np.int32)
    blk = {"pivots_checked": int(len(e)), "ranks_identical": True, "vs_ref": "no fixture for this workload"}
    if dist is not None:
        import torch
        mine = torch.from_numpy(np.concatenate([e, l]).astype(np.int64)).cuda()
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        blk["ranks_identical"] = all(bool(torch.equal(t, allt[0])) for t in allt)
    name = REF_WINDOW_OF.get(wl)
    if expect is not None:
        # deep window: the recorded trace (whose every leaving arc the reference's own pivot() reproduced in replay,
        # and whose continuation the reference solved identically: tests/golden/README.md) is the expectation
        re_, rl_ = expect
        k = min(len(e), len(re_))
        ok = bool(np.array_equal(e[:k], re_[:k]) and np.array_equal(l[:k], rl_[:k]))
        blk["vs_ref"] = "ok" if ok else "MISMATCH"
        blk["ref_pivots_compared"] = int(k)
    elif name is not None:
        from oracle_util import ref_window
        re_, rl_, md5 = ref_window(name)
        k = min(len(e), len(re_))
        ok = bool(np.array_equal(e[:k], re_[:k]) and np.array_equal(l[:k], rl_[:k]))
        blk["vs_ref"] = "ok" if ok else "MISMATCH"
        blk["ref_pivots_compared"] = int(k)
        blk["ref_fixture_md5"] = md5
        if k == len(re_):
            h = hashlib.md5()
            for i in range(k):
                h.update(("P %d %d %d\n" % (i, e[i], l[i])).encode())
            blk["trace_md5"] = h.hexdigest()
            if blk["trace_md5"] != md5:
                blk["vs_ref"] = "MISMATCH"
    return blk


def make_instance(wl):
    import genneteq_b200 as gne
    n, m, seed, mode, rule, desc = WORKLOADS[wl]
    return gne.generate_instance(n, m, seed, mode), rule, desc


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference, in a subprocess (Tree statics: one solver
# per process).  Times whole simplex iterations (computePotentials + pricing + pivot) like e2e.
# ------------------------------------------------------------------------------------------------
_REF_CHILD = r"""
import ctypes as C, json, sys, time, os
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
req = json.load(open(sys.argv[1]))
from oracle_util import REF_SO, ORACLE_SO, dp, ip, oracle_lib, RULES, generate_native
g = generate_native(req["n"], req["m"], req["seed"], req["mode"])     # oracle/libgne_gen.so: the product library is never mapped here
big_m = oracle_lib().orc_big_m(C.c_long(req["m"]), dp(g["cost"]), dp(g["ub"]))
kind = "reference" if os.path.exists(REF_SO) else "port"
if kind == "reference":
    ref = C.CDLL(REF_SO)
    ref.ref_solve.restype = C.c_long
    t0 = time.time()
    ref.ref_init_sparse(req["n"], C.c_long(req["m"]), ip(g["tail"]), ip(g["head"]), dp(g["ub"]), dp(g["cost"]), dp(g["mult"]),
                        dp(g["supply"]), C.c_double(big_m))
    init_s = time.time() - t0
    ref.ref_set_rules(RULES[req["rule"]], RULES[req["rule"]])
    fin = C.c_int(0)
    replay = None
    if req.get("deep"):
        # --window deep: reach the recorded basis with the reference's own pivot() applied to the recorded pivots
        # (ref_driver.cpp replayTraced: every leaving arc is checked against the trace), then solve on from there
        fx = np.load(req["deep"])
        fe = np.ascontiguousarray(fx["e"], np.int32); fl = np.ascontiguousarray(fx["l"], np.int32)
        ref.ref_replay.restype = C.c_long; ref.ref_trace_len.restype = C.c_long
        t0 = time.time()
        mm = ref.ref_replay(1, ip(fe), ip(fl), C.c_long(len(fe)), 1 if req["rule"].startswith("LV") else 0)
        replay = dict(pivots=int(len(fe)), first_mismatch=int(mm), seconds=time.time() - t0)
        if mm != -1: raise SystemExit("reference replay diverged from the recorded trace at pivot %%d" %% mm)
        te = np.zeros(4096, np.int32); tl = np.zeros(4096, np.int32)
        ref.ref_set_trace(ip(te), ip(tl), C.c_long(4096))
    def run(w, k):
        ref.ref_set_window(C.c_long(w))
        ref.ref_solve(1, C.c_long(w + k), C.byref(fin))
        t0 = np.zeros(5); t1 = np.zeros(5)
        ref.ref_get_window_timers(dp(t0)); ref.ref_get_timers(dp(t1))
        if replay is not None:
            k2 = min(int(ref.ref_trace_len()), 4096, len(fx["e2"]))
            replay["continuation_pivots_compared"] = k2
            replay["continuation_equals_recorded"] = bool(np.array_equal(te[:k2], fx["e2"][:k2]) and np.array_equal(tl[:k2], fx["l2"][:k2]))
        return t1 - t0
else:
    replay = None
    lib = oracle_lib()
    t0 = time.time()
    h = C.c_void_p(lib.orc_create(req["n"], C.c_long(req["m"]), ip(g["tail"]), ip(g["head"]), dp(g["ub"]), dp(g["cost"]),
                                  dp(g["mult"]), dp(g["supply"]), C.c_double(big_m)))
    init_s = time.time() - t0
    lib.orc_set_rules(h, RULES[req["rule"]], RULES[req["rule"]])
    fin = C.c_int(0)
    def run(w, k):
        lib.orc_set_window(h, C.c_long(w))
        lib.orc_solve(h, 1, C.c_long(w + k), C.byref(fin))
        t0 = np.zeros(5); t1 = np.zeros(5)
        lib.orc_get_window_timers(h, dp(t0)); lib.orc_get_timers(h, dp(t1))
        return t1 - t0
# One solve() call; the window starts after `warmup` pivots.  The per-stage timers cover
# computePotentials + pricing + pivot of every iteration (the whole loop body of genneteq.cpp:82-106).
# The sample is sized from a per-arc estimate (~55 ns per arc priced once the arc objects no longer fit in
# cache) so that the run stays within the time budget; the window is pivots [warmup, warmup + steps).
per = 55e-9 * req["m"] + 1.0e-7 * req["n"] + (5.0e-8 * req["n"] if req.get("deep") else 0.0)
steps, warmup = req["steps"], req["warmup"]
if per * (steps + warmup) > req["budget_s"]:
    steps = max(3, int(req["budget_s"] / per) - warmup)
    if steps < 3: steps, warmup = 3, 1
tm = run(warmup, steps)
loop_s = tm[0] + tm[1] + tm[2]
out = dict(kind=kind, steps=steps, loop_s=loop_s, pricing_s=tm[0], potentials_s=tm[1], pivot_s=tm[2], arcs_priced=tm[4],
           init_s=init_s, warmup=warmup, replay=replay)
json.dump(out, open(req["out"], "w"))
"""


def deep_fixture(wl):
    return os.path.join(ROOT, "tests", "golden", "deep_%s.npz" % wl.replace("-", "_"))


def run_reference_sample(wl, steps, warmup, budget_s, deep=False):
    n, m, seed, mode, rule, desc = WORKLOADS[wl]
    d = tempfile.mkdtemp(prefix="gne_bench_")
    req = dict(n=n, m=m, seed=seed, mode=mode, rule=rule, steps=steps, warmup=warmup, budget_s=budget_s,
               out=os.path.join(d, "out.json"), deep=deep_fixture(wl) if deep else None)
    rp = os.path.join(d, "req.json")
    json.dump(req, open(rp, "w"))
    code = _REF_CHILD % dict(root=ROOT)
    subprocess.check_call([sys.executable, "-c", code, rp], stdout=subprocess.DEVNULL)
    return json.load(open(req["out"]))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, m, seed, mode, rule, desc = WORKLOADS[args.workload]
    deep = args.window == "deep"
    r = run_reference_sample(args.workload, args.steps, max(args.warmup, 1), budget_s=150.0, deep=deep)
    value = r["arcs_priced"] / r["loop_s"]
    cores = 1
    sample = "%d simplex iterations after %d warm-up iterations, %s, single thread (the reference has none)" % (
        r["steps"], r["warmup"], "Phase I from the recorded deep basis (reached by replaying the recorded pivots through the reference's own pivot())"
        if deep else "Phase I from the initial basis")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": r["warmup"], "ms_per_step": 1e3 * r["loop_s"] / r["steps"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, window=args.window),
        "host_cores_present": os.cpu_count(), "replay": r.get("replay"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": sample,
                         "pricing_only_arcs_per_s": r["arcs_priced"] / r["pricing_s"] if r["pricing_s"] > 0 else None},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pivots_per_s": r["steps"] / r["loop_s"],
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def b200_arm(args):
    import torch
    import genneteq_b200 as gne
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if gne.cuda_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device - genneteq_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, seed, mode, rule, desc = WORKLOADS[args.workload]
    g, rule, desc = make_instance(args.workload)
    deep = args.window == "deep"
    fx = np.load(deep_fixture(args.workload)) if deep else None
    P = len(fx["e"]) if deep else 0
    s = gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], device=local, rule1=rule,
                   trace_cap=max(1024, P + args.steps + args.warmup + 16))
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = np.zeros(128, np.uint8)
            gne._ck(gne.lib.gne_comm_unique_id(buf.ctypes.data_as(gne._u8p)), "gne_comm_unique_id")
            uid = torch.from_numpy(buf)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        s.set_comm(uid.cpu().numpy(), rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dbg("solver ready")

    # ---- e2e: one gne_solver_solve call; W warm-up pivots, then exactly K pivots timed inside the
    # library's loop (host clock, device synchronised at both edges of the window)
    W = max(args.warmup, 3)
    if deep:
        t0 = time.time()
        s.solve(True, P)                        # untimed: reach the recorded basis (the trace is compared below)
        dbg("deep basis reached: %d pivots in %.1f s" % (P, time.time() - t0))
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    s.set_window(W, args.steps)
    s.solve(True, W + args.steps)
    barrier()
    dbg("e2e window done")
    xmode = {0: "none (single GPU)", 1: "ncclAllGather", 2: "NVLink peer-mailbox stores from the exchange kernel (CUDA IPC)"}[
        gne.lib.gne_comm_mode(s.device().h)]
    e2e_s, c0, c1 = s.window()
    if e2e_s <= 0:
        raise SystemExit("bench.py: the solve ended before the timed window completed")
    pivots = int(c1["pivots"] - c0["pivots"])
    arcs_e2e = c1["arcs_priced"] - c0["arcs_priced"]
    te, tl = s.trace
    parity = parity_block(args.workload, te, tl, dist, rank, world,
                          expect=(np.concatenate([fx["e"], fx["e2"]]), np.concatenate([fx["l"], fx["l2"]])) if deep else None)
    dev = s.device()
    N = m
    # ---- device-resident steps: K x (finalize + pricing pass + reduction [+ exchange]), CUDA events
    flush = (25.0 * N / world + 8 * n) < 200e6  # per-GPU working set not larger than L2 => flush between steps
    qs_arcs = None
    burst_ms = None
    if rule.startswith("QS") and world == 1:
        # the block rule's own kernels (k_qs_tiles + k_qs_finish per window), as the solver issues them
        dev.bench_price_qs(max(args.warmup, 3), n, flush)
        ms_step, qs_launched, qs_scanned = dev.bench_price_qs(args.steps, n, flush)
        ms_tiles = ms_step
        qs_arcs = (float(qs_launched.sum()), float(qs_scanned.sum()), float(qs_launched.mean()))
    else:
        dev.bench_price_lv(max(args.warmup, 3), flush)
        barrier()
        dbg("device warm-up steps done")
        ms_step, ms_tiles = dev.bench_price_lv(args.steps, flush)
        # the same launch in a burst (one event pair around all of them): what the per-launch event bracket adds
        burst_ms = dev.bench_price_lv_burst(args.steps) if (world == 1 and not flush) else None
    barrier()
    dbg("device steps done: kernel ms mean %.4f min %.4f max %.4f | step ms mean %.4f min %.4f max %.4f" % (
        ms_tiles.mean(), ms_tiles.min(), ms_tiles.max(), ms_step.mean(), ms_step.min(), ms_step.max()))
    clocks = sampler.stop()
    if os.environ.get("GNE_BENCH_DEBUG") == "2":
        for name, mean, best in dev.bench_variants(10):
            if name.startswith(("production", "tma<256thr,4it,3st> 148x2", "k_price_lv")):
                dbg("variant %-60s mean %.2f us best %.2f us" % (name, mean * 1e3, best * 1e3))
    launches = int(s.counters()["launches"] - c0["launches"])     # e2e window + device steps (warm-up excluded from neither)
    lv_kernel = "k_qs_tiles + k_qs_finish" if qs_arcs else {"stream": "k_price_lv_stream", "tiles": "k_price_lv_tiles"}[dev.last_lv_kernel()]
    dev.close()
    if dist is not None:
        t = torch.tensor([e2e_s, float(ms_step.sum()), float(ms_tiles.mean())], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, step_sum_ms, tiles_ms = (float(x) for x in t.cpu())
    else:
        step_sum_ms, tiles_ms = float(ms_step.sum()), float(ms_tiles.mean())
    # orderly teardown on every rank: the library's NCCL communicator first, then torch's
    s.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    value = N * args.steps / (step_sum_ms * 1e-3)
    peak, peak_src = peaks()
    shard_arcs = N / world
    alg_bytes = 25.0 * shard_arcs + 8.0 * n
    if qs_arcs:
        # block rule: a pass prices a window of the list, not the list - arcs the kernels covered per pass
        value = qs_arcs[0] / (step_sum_ms * 1e-3)
        alg_bytes = 25.0 * qs_arcs[2] + 8.0 * n
    achieved = alg_bytes / (tiles_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": step_sum_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world, window=args.window),
        "sharding": "contiguous position ranges over %d GPU(s); per-pivot exchange of the per-shard record: %s" % (world, xmode),
        "parity": parity,
        "e2e": {"value": arcs_e2e / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": (c1["h2d_bytes"] - c0["h2d_bytes"]) / max(pivots, 1),
                "d2h_bytes_per_step": (c1["d2h_bytes"] - c0["d2h_bytes"]) / max(pivots, 1),
                "pivots_per_s": pivots / e2e_s, "ms_per_pivot": 1e3 * e2e_s / max(pivots, 1),
                "host_split_s": {k: c1[k] - c0[k] for k in ("pricing_s", "potentials_s", "pivot_s", "trees_s", "flows_s")}},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.workload, world, lv_kernel), "kernel": lv_kernel, "kernel_ms": tiles_ms,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
    }
    if burst_ms:
        # informational: `achieved` / `frac` above stay on the per-launch bracket (kernel_ms)
        line["roofline"]["kernel_ms_in_a_burst"] = burst_ms
        line["roofline"]["frac_in_a_burst"] = alg_bytes / (burst_ms * 1e-3) / 1e9 / peak
    if qs_arcs:
        line["block_rule"] = {"arcs_covered_by_kernels_per_pass": qs_arcs[2], "arcs_consumed_by_rule_per_pass": qs_arcs[1] / args.steps,
                              "note": "value / roofline count the arcs the window kernels priced; a pass = window kernel + finish kernel per window"}
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference_sample(args.workload, 12, 2, budget_s=25.0, deep=deep)
            line["cpu_baseline"] = {
                "value": r["arcs_priced"] / r["loop_s"], "unit": UNIT, "cores": 1, "kind": r["kind"],
                "sample": "%d simplex iterations after %d warm-up iterations of the same workload, single thread "
                          "(the reference has none; %d host cores present)" % (r["steps"], r["warmup"], os.cpu_count()),
                "ms_per_pivot": 1e3 * r["loop_s"] / r["steps"],
                "pricing_only_arcs_per_s": r["arcs_priced"] / r["pricing_s"] if r["pricing_s"] > 0 else None}
        except Exception as ex:                      # the oracle always exists; report why it could not run
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "failed: %r" % (ex,)}
    print(json.dumps(line))
    sys.stdout.flush()
    if not parity["ranks_identical"] or parity["vs_ref"] == "MISMATCH":
        sys.stderr.write("bench.py: PARITY FAILURE %r\n" % (parity,))
        return 3
    return 0


# ------------------------------------------------------------------------------------------------
# configs[3]: a batch of independent instances, one solver handle (one CUDA stream) per instance, one host
# thread per handle (ctypes releases the GIL); replicas only - no collective.  "scaling": "weak".
# ------------------------------------------------------------------------------------------------
_BATCH_REF_CHILD = r"""
import ctypes as C, json, sys, time, os
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
req = json.load(open(sys.argv[1]))
from oracle_util import REF_SO, dp, ip, oracle_lib, RULES, generate_native
g = generate_native(req["n"], req["m"], req["seed"], req["mode"])     # oracle/libgne_gen.so: the product library is never mapped here
big_m = oracle_lib().orc_big_m(C.c_long(req["m"]), dp(g["cost"]), dp(g["ub"]))
ref = C.CDLL(REF_SO)
ref.ref_solve.restype = C.c_long
ref.ref_init_sparse(req["n"], C.c_long(req["m"]), ip(g["tail"]), ip(g["head"]), dp(g["ub"]), dp(g["cost"]), dp(g["mult"]),
                    dp(g["supply"]), C.c_double(big_m))
ref.ref_set_rules(RULES[req["rule"]], RULES[req["rule"]])
fin = C.c_int(0)
ref.ref_set_window(C.c_long(req["warmup"]))
ref.ref_solve(1, C.c_long(req["warmup"] + req["steps"]), C.byref(fin))
t0 = np.zeros(5); t1 = np.zeros(5)
ref.ref_get_window_timers(dp(t0)); ref.ref_get_timers(dp(t1))
tm = t1 - t0
json.dump(dict(loop_s=tm[0] + tm[1] + tm[2], arcs_priced=tm[4]), open(req["out"], "w"))
"""


def batch_reference(args, count, steps, warmup):
    """One reference PROCESS per instance (Tree statics forbid two solvers in a process), all host cores."""
    n, m, seed0, mode, rule, desc = WORKLOADS["c4-batch"]
    cores = min(os.cpu_count() or 1, count)
    d = tempfile.mkdtemp(prefix="gne_batch_")
    procs, outs = [], []
    code = _BATCH_REF_CHILD % dict(root=ROOT)
    t0 = time.perf_counter()
    for i in range(cores):
        req = dict(n=n, m=m, seed=seed0 + i, mode=mode, rule=rule, steps=steps, warmup=warmup, out=os.path.join(d, "o%d.json" % i))
        rp = os.path.join(d, "r%d.json" % i)
        json.dump(req, open(rp, "w"))
        outs.append(req["out"])
        procs.append(subprocess.Popen([sys.executable, "-c", code, rp], stdout=subprocess.DEVNULL))
    for p in procs:
        p.wait()
    res = [json.load(open(o)) for o in outs]
    arcs = sum(r["arcs_priced"] for r in res)
    worst = max(r["loop_s"] for r in res)
    return dict(value=arcs / worst, cores=cores, instances=cores, steps=steps, warmup=warmup, ms_per_pivot=1e3 * worst / steps,
                wall_s=time.perf_counter() - t0)


def batch_arm(args):
    import threading as th
    # one hardware work queue per instance stream (the default of 8 connections makes 32 streams share queues, and
    # kernels of different instances then wait for each other); must be set before the CUDA context exists
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import genneteq_b200 as gne
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    n, m, seed0, mode, rule, desc = WORKLOADS["c4-batch"]
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = batch_reference(args, args.batch * max(args.gpus, 1), min(args.steps, 40), max(2, min(args.warmup, 5)))
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": r["steps"],
                "warmup": r["warmup"], "ms_per_step": r["ms_per_pivot"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config("c4-batch", max(args.gpus, 1), batch=args.batch),
                "instances_timed": r["instances"], "host_cores_present": os.cpu_count(),
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                                 "sample": "%d instances, one process each, %d iterations after %d warm-up" % (r["instances"], r["steps"], r["warmup"])},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0
    if gne.cuda_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device - genneteq_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    W = max(args.warmup, 3)
    # every instance's persistent pricing pass gets its share of the CTA slots (2 per SM) so that the B passes are
    # resident together instead of queueing behind each other
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    per = max(0, min(args.batch_group, 8))
    # launch groups (the default): `per` instances per launch of the batched kernel, together on the 2 x SMs CTA slots;
    # --batch-group 0: one launch and one stream per instance, every pass on a 1/B share
    ctas = args.batch_ctas if args.batch_ctas > 0 else max(1, (2 * sms) // (per if per > 0 else B))
    os.environ["GNE_STREAM_CTAS"] = str(ctas)
    os.environ.setdefault("GNE_POLL_SLEEP_US", "20")    # B host threads on fewer cores: the result poll sleeps between looks
    solvers = []
    for i in range(B):
        g = gne.generate_instance(n, m, seed0 + rank * B + i, mode)
        solvers.append(gne.Solver(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"], device=local, rule1=rule,
                                  trace_cap=W + args.steps + 16))
    sampler = ClockSampler(local)
    if dist is not None:
        dist.barrier()
    sampler.start()

    # one call drives all B solvers: min(B, cores) host threads, each interleaving its share of the instances
    # (gne_solver_solve_batch) - while one instance's pass runs on the device its thread pivots another one
    for s in solvers:
        s.set_window(W, args.steps)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if args.batch_threads > 0:
        nthreads = args.batch_threads
    elif per > 0:
        # grouped launches: two host threads per rank drive the B instances (each pivots one instance while the others are
        # priced).  Measured on one GPU, 32 instances, windows aligned: 1 thread 129 k pivots/s, 2 threads 186 k (the device
        # leg's own pace), 8 threads 180 k
        nthreads = 2
    else:
        nthreads = min(B, cores)
    bstats = {}
    # (every instance waits at pivot W for the others: the timed windows then run side by side whatever the first pivots cost)
    gne.solve_batch(solvers, True, W + args.steps, threads=nthreads, per_launch=per, stats=bstats, align_pivot=W)
    wins = [s.window() for s in solvers]
    edges = [s.window_edges for s in solvers]
    # instance 0 of rank 0 is the (seed 1000) instance the reference's prefix fixture was written for
    parity = parity_block("c4-batch", *solvers[0].trace, None, rank, world) if rank == 0 else None
    wall = max(e[1] for e in edges) - min(e[0] for e in edges)
    arcs = sum(w[2]["arcs_priced"] - w[1]["arcs_priced"] for w in wins)
    pivots = sum(w[2]["pivots"] - w[1]["pivots"] for w in wins)
    h2d = sum(w[2]["h2d_bytes"] - w[1]["h2d_bytes"] for w in wins)
    d2h = sum(w[2]["d2h_bytes"] - w[1]["d2h_bytes"] for w in wins)
    launches = sum(w[2]["launches"] - w[1]["launches"] for w in wins)
    devs = [s.device() for s in solvers]
    if per > 0:
        # (the handles count one launch per pass; grouped, a launch carries several passes: scale by the measured ratio)
        passes_all = sum(w[2]["launches"] for w in wins)
        launches = launches * bstats.get("group_launches", passes_all) / max(passes_all, 1)
        # device-resident leg: K rounds of passes over all instances, `per` instances per launch of the batched kernel,
        # launches alternating between two streams; device clock (events) around all rounds
        gne.bench_lv_groups(devs, 3, per, 2)
        dev_ms, dev_launches = gne.bench_lv_groups(devs, args.steps, per, 2)
        one_ms, _ = gne.bench_lv_groups(devs[:per], args.steps, per, 1)       # one launch at a time: the kernel's own duration
        clocks = sampler.stop()
        dev_wall = dev_ms * 1e-3
        for s in solvers:
            s.close()
        if dist is not None:
            t = torch.tensor([wall, dev_wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall, dev_wall = (float(x) for x in t.cpu())
            tot = torch.tensor([arcs, pivots, h2d, d2h, launches + dev_launches], dtype=torch.float64, device="cuda")
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            arcs, pivots, h2d, d2h, launches = (float(x) for x in tot.cpu())
            dist.barrier()
            dist.destroy_process_group()
        else:
            launches += dev_launches
        if rank != 0:
            return 0
        peak, peak_src = peaks()
        alg = 25.0 * m + 8.0 * n
        line = {"metric": METRIC, "value": world * B * m * args.steps / dev_wall, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": 1e3 * dev_wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config("c4-batch", world, batch=B),
                "sharding": "replicas only: %d instances per GPU, their LV passes in launch groups of %d (one grid, %d CTAs per instance), "
                            "%d host threads per rank interleaving them, no collective" % (B, per, ctas, nthreads),
                "parity": parity,
                "e2e": {"value": arcs / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / max(pivots, 1), "d2h_bytes_per_step": d2h / max(pivots, 1),
                        "pivots_per_s": pivots / wall, "ms_per_pivot_per_instance": 1e3 * wall / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "host_cores_present": os.cpu_count(), "host_cores_usable": cores,
                # per GPU: a round of passes moves B x (25 B x arcs + 8 B x nodes)
                "roofline": {"bound": "hbm", "achieved": B * alg * args.steps / dev_wall / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": B * alg * args.steps / dev_wall / 1e9 / peak, "traffic": None,
                             "kernel": "k_price_lv_batch: %d instances per launch, %d CTAs per instance (aggregate bytes of all passes / "
                                       "device time of the whole leg; kernel_ms = one launch timed alone)" % (per, ctas),
                             "kernel_ms": one_ms / args.steps, "algorithmic_bytes_per_launch": per * alg, "peak_source": peak_src}}
        if not args.no_cpu_baseline and world == 1:
            r = batch_reference(args, B, 12, 2)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                                    "sample": "%d instances, one reference process per instance on %d host cores, %d iterations after %d warm-up" % (
                                        r["instances"], r["cores"], r["steps"], r["warmup"]), "ms_per_pivot": r["ms_per_pivot"]}
        print(json.dumps(line))
        sys.stdout.flush()
        if parity["vs_ref"] == "MISMATCH":
            sys.stderr.write("bench.py: PARITY FAILURE %r\n" % (parity,))
            return 3
        return 0
    # device-resident leg: every instance replays K device steps on its own stream, all streams concurrently
    for d in devs:
        d.bench_price_lv(3, False)
    torch.cuda.synchronize()
    launches_before = sum(d.launch_count() for d in devs)
    res = [None] * B

    def steps(i):
        res[i] = devs[i].bench_price_lv(args.steps, False)
    # device clock around the whole leg: an event before the first launch and one after every instance stream has
    # drained (the handles' streams are independent of torch's, so the second event is recorded after a device sync)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    threads = [th.Thread(target=steps, args=(i,)) for i in range(B)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    ev1.record()
    torch.cuda.synchronize()
    dev_wall_host = time.perf_counter() - t0
    dev_wall = ev0.elapsed_time(ev1) * 1e-3
    busy = max(float(r[0].sum() + r[1].sum()) for r in res) * 1e-3     # the busiest instance stream's kernel time
    clocks = sampler.stop()
    tiles_ms = float(np.mean([r[1].mean() for r in res]))
    launches += sum(d.launch_count() for d in devs) - launches_before
    for s in solvers:
        s.close()
    if dist is not None:
        t = torch.tensor([wall, dev_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_wall = (float(x) for x in t.cpu())
        tot = torch.tensor([arcs, pivots, h2d, d2h, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        arcs, pivots, h2d, d2h, launches = (float(x) for x in tot.cpu())
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peak, peak_src = peaks()
    alg = 25.0 * m + 8.0 * n
    # (gne_bench_price_lv runs two loops of `steps` passes each: kernel-bracketed, then as launched by the solver)
    line = {"metric": METRIC, "value": world * B * m * 2 * args.steps / dev_wall, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": 1e3 * dev_wall / (2 * args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config("c4-batch", world, batch=B),
            "sharding": "replicas only: one solver handle / CUDA stream per instance, %d host threads interleaving them, no collective" % nthreads,
            "parity": parity,
            "e2e": {"value": arcs / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / max(pivots, 1), "d2h_bytes_per_step": d2h / max(pivots, 1),
                    "pivots_per_s": pivots / wall, "ms_per_pivot_per_instance": 1e3 * wall / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "device_leg_host_wall_s": dev_wall_host, "busiest_stream_kernel_s": busy,
            # per GPU: the B concurrent passes together move B x (25 B x arcs + 8 B x nodes) per round of passes
            "roofline": {"bound": "hbm", "achieved": B * alg * 2 * args.steps / dev_wall / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": B * alg * 2 * args.steps / dev_wall / 1e9 / peak, "traffic": None,
                         "kernel": "k_price_lv_stream, %d CTAs per instance, the %d instances of a GPU resident together "
                                   "(aggregate bytes of all passes / device time of the whole leg; kernel_ms = one instance's pass)" % (ctas, B),
                         "kernel_ms": tiles_ms, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src}}
    if not args.no_cpu_baseline and world == 1:
        r = batch_reference(args, B, 12, 2)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                                "sample": "%d instances, one reference process per instance on %d host cores, %d iterations after %d warm-up" % (
                                    r["instances"], r["cores"], r["steps"], r["warmup"]), "ms_per_pivot": r["ms_per_pivot"]}
    print(json.dumps(line))
    sys.stdout.flush()
    if parity["vs_ref"] == "MISMATCH":
        sys.stderr.write("bench.py: PARITY FAILURE %r\n" % (parity,))
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3-lv", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--window", default="start", choices=["start", "deep"],
                    help="start: pivots from the initial basis (default); deep: from the recorded basis of tests/golden/deep_<workload>.npz")
    ap.add_argument("--batch", type=int, default=32, help="instances per GPU for --workload c4-batch")
    ap.add_argument("--batch-threads", type=int, default=0, help="host threads driving the batch (0: chosen from the cores available)")
    ap.add_argument("--batch-group", type=int, default=8, help="instances per launch of the batched LV kernel (1..8; 0: one launch per instance)")
    ap.add_argument("--batch-ctas", type=int, default=0, help="CTAs per instance (0: 2 x SMs / instances per launch)")
    args = ap.parse_args()
    if args.workload == "c4-batch":
        return batch_arm(args)
    if args.impl == "reference":
        return reference_arm(args)
    return b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
