"""N > 1 path.  CPU: world_size-2 gloo test of the host merge logic.  GPU: the sharded solver on every GPU of the box -
one rank per GPU bootstrapped through NCCL when the box has >= 2 GPUs, otherwise (single-GPU box) two or more ranks
SHARING device 0, bootstrapped through torch's gloo group with the library's all-gather supplied as a host callback
(gne_solver_set_comm_host; NCCL refuses two ranks on one device).  The per-pivot exchange is tested both ways in either
setting: through the CUDA-IPC peer mailboxes written from the pricing / exchange kernels, and through the fallback
(ncclAllGather, or the host callback).  No test here skips on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def launch(mode, nproc, out, port, extra_env=None, rule="LV"):
    env = dict(os.environ)
    env.update(extra_env or {})
    env["MASTER_ADDR"] = "127.0.0.1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "mgpu_worker.py"), mode, out, rule]
    r = subprocess.run(cmd, env=env, timeout=1500, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    return open(out).read()


def ranks_for_box(gne, want=None):
    """Ranks to launch: every GPU of the box (2, 4 or 8), or 2 ranks sharing the one GPU of a single-GPU box."""
    ngpu = gne.cuda_device_count()
    assert ngpu >= 1, "no CUDA device"
    if want is not None:
        return want
    return 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)


def fallback_env(gne, nproc):
    """Environment that forces the non-mailbox exchange: NCCL with one rank per GPU, else the host callback."""
    return {"GNE_EXCHANGE": "nccl"} if gne.cuda_device_count() >= nproc else {"GNE_EXCHANGE": "host"}


def test_sharded_lv_exchange_host_logic_gloo(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    assert launch("gloo", 2, str(tmp_path / "gloo.txt"), 29611) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("rule,exchange", [("QS", "p2p"), ("CL", "p2p"), ("QS", "fallback")])
def test_sharded_solve_two_ranks_other_rules(tmp_path, rule, exchange):
    """quickSelect (blob exchange) and the candidate-list rule (gathered violator lists) over 2 ranks: full C1 solve,
    trace / objective / flows equal the reference's on every rank."""
    import genneteq_b200 as gne
    env = fallback_env(gne, 2) if exchange == "fallback" else {}
    port = 29620 + ["QS", "CL"].index(rule) * 2 + (1 if exchange == "fallback" else 0)
    assert launch("nccl", 2, str(tmp_path / (rule + exchange + ".txt")), port, env, rule) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "fallback"])
def test_sharded_solve_two_ranks(tmp_path, exchange):
    """Full C1 solve sharded over 2 ranks: the per-pivot exchange through peer mailboxes (default) and through the
    fallback (ncclAllGather / host callback) must both reproduce the reference's pivot trace."""
    import genneteq_b200 as gne
    env = fallback_env(gne, 2) if exchange == "fallback" else {}
    port = 29612 if exchange == "p2p" else 29613
    assert launch("nccl", 2, str(tmp_path / (exchange + ".txt")), port, env) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("rule", ["LV", "QS"])
def test_sharded_c2_prefix_window_all_ranks(tmp_path, rule):
    """configs[1] size (100k nodes / 2M arcs) sharded over every GPU of the box (or 2 ranks on one GPU): the first 300
    pivots on every rank equal the prefix written by the UNMODIFIED reference (tests/golden/ref_windows.npz), flows
    and potentials equal the oracle's, and every rank holds the same trace."""
    import genneteq_b200 as gne
    nproc = ranks_for_box(gne)
    assert launch("nccl-c2", nproc, str(tmp_path / ("c2" + rule + ".txt")), 29640 + (1 if rule == "QS" else 0), None, rule) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nccl", "nccl-c2"])
def test_sharded_streamed_pass_with_fused_exchange(tmp_path, mode):
    """The TMA-streamed pass whose last CTA runs the reduction AND the peer-mailbox exchange (the path configs[2] takes
    on 2-8 GPUs), forced onto small shards with GNE_LV_KERNEL=stream: the full C1 solve (long tie lists, overflow
    fallbacks, both phases) on 2 ranks and the configs[1] prefix window on every rank must reproduce the reference trace."""
    import genneteq_b200 as gne
    nproc = 2 if mode == "nccl" else ranks_for_box(gne)
    port = 29660 + (1 if mode == "nccl-c2" else 0)
    assert launch(mode, nproc, str(tmp_path / ("fused" + mode + ".txt")), port, {"GNE_LV_KERNEL": "stream"}, "LV") == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,exchange", [(2, "p2p"), (8, "fallback")])
def test_sharded_lists_shorter_than_the_shard_grid(tmp_path, nproc, exchange):
    """Lists shorter than 2048 x ranks leave some ranks with an EMPTY shard (ranges must never overlap): the edge
    fixtures (2- and 3-node graphs, infeasible demand, zero supplies) with LV / QS / CL over 2 ranks, and C1 (10k arcs:
    5 non-empty shards) over 8 ranks, reproduce the reference on every rank."""
    import genneteq_b200 as gne
    env = fallback_env(gne, nproc) if exchange == "fallback" else {}
    if nproc == 2:
        assert launch("edge", 2, str(tmp_path / "edge.txt"), 29670, env) == "OK"
    else:
        if gne.cuda_device_count() < nproc:
            env["GNE_TEST_BOOT"] = "host"
        assert launch("nccl", nproc, str(tmp_path / "c1x8.txt"), 29671, env, "LV") == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,exchange", [(2, "p2p"), (4, "fallback")])
def test_sharded_first_violator_and_candidate_list_2(tmp_path, nproc, exchange):
    """gnePivotRules.cpp:206-228 / :613-677 over a sharded list: FV = the lowest of the shards' first violating positions,
    CL2 / CLW2 = the shards' parts of the scanned range concatenated in cursor order, RV through the gathered violator
    lists; 600 nodes / 9000 arcs (several non-empty shards) must reproduce the reference on every rank."""
    import genneteq_b200 as gne
    env = fallback_env(gne, nproc) if exchange == "fallback" else {}
    if gne.cuda_device_count() < nproc:
        env["GNE_TEST_BOOT"] = "host"
    assert launch("medium", nproc, str(tmp_path / "medium.txt"), 29680 + nproc, env, "FV,CL2,CLW2,RV") == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,exchange", [(2, "p2p"), (3, "fallback")])
def test_sharded_equal_flow_instance(tmp_path, nproc, exchange):
    """An instance with equal-flow sets sharded over the ranks (gne_solver_set_comm[_host] on the equal-flow engine): arcs priced
    per shard and merged, sets priced on every rank; Dantzig, block rule, candidate lists 1 and 2 - every rank
    the reference's trace, flows and potentials (fixture q_pow2_s29, written by the reference)."""
    gne = pytest.importorskip("genneteq_b200")
    if gne.cuda_device_count() < 1:
        pytest.skip("needs a GPU")
    env = fallback_env(gne, nproc) if exchange == "fallback" else {}
    assert launch("eqf", nproc, str(tmp_path / "eqf.txt"), 29690 + nproc, env, "LV,QS,CLW2" if exchange == "p2p" else "LV,CL,CL2") == "OK"
