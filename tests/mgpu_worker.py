"""Worker for the multi-rank tests (launched with torch.distributed.run).

mode "nccl": one rank per GPU; the solver prices its shard of the nonbasic list on its GPU and the
             per-shard records are exchanged with NCCL inside the library.  Every rank must produce
             the reference's pivot trace.
mode "gloo": CPU only; each rank prices its position range with the oracle's scalar pass, the
             records are all-gathered with gloo and merged by gne_merge_lv_shards (host logic of the
             N > 1 path).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(HERE, "golden")):
    sys.path.insert(0, p)

import genneteq_b200 as gne  # noqa: E402
from oracle_util import RULES, Instance, Oracle, oracle_lib, dp, ip, lp, bp, ref_window  # noqa: E402


class Boot:
    """Process group + the way the solver is attached to it.  One rank per GPU with NCCL when the box has enough
    GPUs; otherwise (single-GPU box) the ranks share device 0, torch's gloo group carries the bootstrap and the
    library gets its all-gather as a host callback (gne_solver_set_comm_host) - NCCL refuses two ranks on one device."""

    def __init__(self, rank, world, local):
        self.rank, self.world = rank, world
        ngpu = torch.cuda.device_count()
        self.shared = ngpu < world or os.environ.get("GNE_TEST_BOOT") == "host"
        self.device = local % max(ngpu, 1)
        torch.cuda.set_device(self.device)
        if self.shared:
            dist.init_process_group("gloo")
            self.tdev = "cpu"
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.device))
            self.tdev = "cuda"

    def _allgather(self, send, nbytes):
        mine = torch.frombuffer(bytearray(send), dtype=torch.uint8)
        outs = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(self.world)]
        dist.all_gather(outs, mine)
        return b"".join(o.numpy().tobytes() for o in outs)

    def attach(self, s):
        if self.shared:
            s.set_comm_host(self.rank, self.world, self._allgather)
            return
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = np.zeros(128, np.uint8)
            gne._ck(gne.lib.gne_comm_unique_id(buf.ctypes.data_as(gne._u8p)), "gne_comm_unique_id")
            uid = torch.from_numpy(buf)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        s.set_comm(uid.cpu().numpy(), self.rank, self.world)

    def same_everywhere(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr).astype(np.int64)).to(self.tdev)
        ref = t.clone()
        dist.broadcast(ref, 0)
        return bool(torch.equal(t, ref))

    def all_ok(self, ok):
        flag = torch.tensor([1 if ok else 0], device=self.tdev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return int(flag.item()) == 1

    def finish(self, out, ok, note=""):
        good = self.all_ok(ok)
        if self.rank == 0:
            open(out, "w").write("OK" if good else "MISMATCH " + note)
        dist.barrier()
        dist.destroy_process_group()


def main():
    mode, out = sys.argv[1], sys.argv[2]
    rule = sys.argv[3] if len(sys.argv) > 3 else "LV"
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    gold = np.load(os.path.join(HERE, "golden", "c1_wide_s7.npz"))
    inst = Instance.load(os.path.join(HERE, "golden", "c1_wide_s7_instance.npz"))
    if mode == "nccl-c2":
        # configs[1] size, sharded: the first 300 pivots on EVERY rank against the prefix the unmodified reference
        # wrote (tests/golden/ref_windows.npz); flows and potentials against the oracle on rank 0
        boot = Boot(rank, world, local)
        g = gne.generate_instance(100_000, 2_000_000, 21, "lossy")
        big = Instance(g["n"], g["tail"], g["head"], g["ub"], g["cost"], g["mult"], g["supply"])
        s = gne.Solver(big.n, big.tail, big.head, big.ub, big.cost, big.mult, big.supply, device=boot.device, rule1=rule)
        boot.attach(s)
        s.solve(True, 300)
        e, l = s.trace
        re_, rl_, _ = ref_window("c2_" + rule.lower())
        ok = len(e) == 300 and np.array_equal(e, re_[:300]) and np.array_equal(l, rl_[:300])
        ok = ok and boot.same_everywhere(np.concatenate([e, l]))
        if rank == 0:
            o = Oracle(big, RULES[rule])
            o.solve(True, 300)
            ok = ok and np.array_equal(e, o.trace[0]) and np.array_equal(l, o.trace[1]) and np.array_equal(s.flows(), o.flows())
            ok = ok and np.array_equal(s.potentials(), o.potentials())
        mode_id = gne.lib.gne_comm_mode(s.device().h)
        s.close()
        boot.finish(out, ok, "comm mode %d" % mode_id)
        return
    if mode == "nccl":
        boot = Boot(rank, world, local)
        s = gne.Solver(inst.n, inst.tail, inst.head, inst.ub, inst.cost, inst.mult, inst.supply, device=boot.device, rule1=rule)
        boot.attach(s)
        want = os.environ.get("GNE_TEST_EXPECT_MODE")
        p1, p2 = s.solve_both()
        mode_id = gne.lib.gne_comm_mode(s.device().h)
        e, l = s.trace
        ok = (p1, p2) == tuple(int(v) for v in gold[rule + "_p"]) and np.array_equal(e, gold[rule + "_e"]) and np.array_equal(l, gold[rule + "_l"])
        ok = ok and s.objective == float(gold[rule + "_objective"]) and np.array_equal(s.flows(), gold[rule + "_flows"])
        if want:
            ok = ok and mode_id in [int(x) for x in want.split(",")]
        s.close()
        boot.finish(out, ok, "comm mode %d" % mode_id)
        return
    if mode == "eqf":
        # an instance WITH equal-flow sets (600 nodes / 9000 arcs / 10 sets: several 2048-position tiles) over the ranks: every
        # rank runs the equal-flow host engine, prices its range of the arcs (LV / QS / first violator / compaction, merged by
        # the device layer) and all of the sets; fixture written by the reference (tests/golden/q_pow2_s29.npz)
        boot = Boot(rank, world, local)
        z = np.load(os.path.join(HERE, "golden", "q_pow2_s29.npz"))
        ok = True
        note = ""
        for r in rule.split(","):
            s = gne.Solver(int(z["n"]), z["tail"], z["head"], z["ub"], z["cost"], z["mult"], z["supply"], device=boot.device, rule1=r,
                           eqf=z["eqf"], num_sets=int(z["p"]))
            boot.attach(s)
            p1, p2 = s.solve_both()
            e, l = s.trace
            good = (p1, p2) == tuple(int(v) for v in z[r + "_p"]) and np.array_equal(e, z[r + "_e"]) and np.array_equal(l, z[r + "_l"])
            good = good and np.array_equal(s.flows(), z[r + "_flows"], equal_nan=True)
            good = good and np.array_equal(s.potentials(), z[r + "_potentials"], equal_nan=True)
            good = good and boot.same_everywhere(np.concatenate([e, l]))
            if not good:
                note += " %s(%d+%d pivots)" % (r, p1, p2)
            ok = ok and good
            s.close()
        boot.finish(out, ok, note)
        return
    if mode == "medium":
        # 600 nodes / 9000 arcs over the ranks (more than one non-empty shard): sharded first-violator (min over the shards'
        # first positions), sharded candidate-list-2 (cursor-ordered concatenation of the shards' parts), random rules
        boot = Boot(rank, world, local)
        from test_oracle_pinned import load
        z, mi = load("m_wide_s9")
        ok = True
        note = ""
        for r in rule.split(","):
            s = gne.Solver(mi.n, mi.tail, mi.head, mi.ub, mi.cost, mi.mult, mi.supply, device=boot.device, rule1=r)
            boot.attach(s)
            p1, p2 = s.solve_both()
            e, l = s.trace
            good = (p1, p2) == tuple(int(v) for v in z[r + "_p"]) and np.array_equal(e, z[r + "_e"]) and np.array_equal(l, z[r + "_l"])
            good = good and np.array_equal(s.flows(), z[r + "_flows"]) and s.objective == float(z[r + "_objective"])
            if not good:
                note += " " + r
            ok = ok and good
            s.close()
        boot.finish(out, ok, note)
        return
    if mode == "edge":
        # lists shorter than 2048 x ranks: some ranks own EMPTY shards; every edge fixture, rules LV / QS / CL
        boot = Boot(rank, world, local)
        ok = True
        note = ""
        for name in ("e_infeasible_s11", "e_nosupply_s12", "e_two_nodes_s13", "e_three_nodes_s14", "s_wide_s3"):
            from test_oracle_pinned import load
            z, ei = load(name)
            for r in ("LV", "QS", "CL"):
                s = gne.Solver(ei.n, ei.tail, ei.head, ei.ub, ei.cost, ei.mult, ei.supply, device=boot.device, rule1=r)
                boot.attach(s)
                p1, p2 = s.solve_both()
                e, l = s.trace
                good = np.array_equal(e, z[r + "_e"]) and np.array_equal(l, z[r + "_l"]) and np.array_equal(s.flows(), z[r + "_flows"])
                if not good:
                    note += " %s/%s" % (name, r)
                ok = ok and good
                s.close()
        boot.finish(out, ok, note)
        return
    # ---- gloo: host logic of the sharded LV exchange
    dist.init_process_group("gloo")
    lib = oracle_lib()
    band = 4.0e-13
    stride = 48
    good = True
    for K in (0, 40, 900):
        o = Oracle(inst, RULES["LV"])
        o.solve(True, K)
        pi = o.compute_potentials()
        nb = o.nonbasic_soa()
        N = len(nb["order"])
        per = ((N + world - 1) // world + 2047) // 2048 * 2048
        lo, hi = rank * per, min(N, (rank + 1) * per)
        sl = slice(lo, max(lo, hi))
        cnt_cap = max(hi - lo, 1)
        pos = np.zeros(cnt_cap, np.int64); viol = np.zeros(cnt_cap)
        t, h, c, m, st = (np.ascontiguousarray(nb[k][sl]) for k in ("tail", "head", "cost", "mult", "state"))
        k = lib.orc_price_violators(C.c_long(max(hi - lo, 0)), ip(t), ip(h), dp(c), dp(m), bp(st), dp(pi), C.c_double(0.0),
                                    lp(pos), dp(viol), C.c_long(cnt_cap))
        pos = pos[:k] + lo; viol = viol[:k]
        rec = np.zeros(3 + 2 * stride)
        if k:
            mx = viol.max()
            keep = viol >= mx - band
            c = int(keep.sum())
            rec[0] = mx; rec[1] = 1; rec[2] = c if c <= stride else -1
            rec[3:3 + min(c, stride)] = pos[keep][:stride]
            rec[3 + stride:3 + stride + min(c, stride)] = viol[keep][:stride]
        else:
            rec[0] = -1.0
        mine = torch.from_numpy(rec)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = np.stack([a.numpy() for a in allr])
        res = gne.merge_lv_shards(allr[:, 0], allr[:, 1].astype(np.int32), allr[:, 2].astype(np.int64),
                                  allr[:, 3:3 + stride].astype(np.int64).ravel(), allr[:, 3 + stride:].ravel(), stride)
        lst = np.zeros(N, np.int64); L = C.c_double(0); anyv = C.c_int(0)
        kk = lib.orc_price_lv(C.c_long(N), ip(nb["tail"]), ip(nb["head"]), dp(nb["cost"]), dp(nb["mult"]), bp(nb["state"]),
                              dp(pi), lp(lst), C.c_long(N), C.byref(L), C.byref(anyv))
        good = good and res["any"] == bool(anyv.value) and not res["need_fallback"]
        good = good and res["largest"] == L.value and np.array_equal(res["list"], lst[:kk])
        o.close()
    flag = torch.tensor([1 if good else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out, "w").write("OK" if int(flag.item()) == 1 else "MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
