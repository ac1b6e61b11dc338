"""N > 1 path: world_size-2 gloo test of the host logic on CPU, NCCL test on >= 2 GPUs."""
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
    subprocess.run(cmd, check=True, env=env, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return open(out).read()


def test_sharded_lv_exchange_host_logic_gloo(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    assert launch("gloo", 2, str(tmp_path / "gloo.txt"), 29611) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("rule,exchange", [("QS", "p2p"), ("CL", "p2p"), ("QS", "nccl")])
def test_sharded_solve_two_gpus_other_rules(tmp_path, rule, exchange):
    """quickSelect (two-round blob exchange) and the candidate-list rule (gathered violator lists) over 2 GPUs."""
    import genneteq_b200 as gne
    if gne.cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = {"GNE_EXCHANGE": "nccl"} if exchange == "nccl" else {}
    port = 29620 + ["QS", "CL"].index(rule) * 2 + (1 if exchange == "nccl" else 0)
    assert launch("nccl", 2, str(tmp_path / (rule + exchange + ".txt")), port, env, rule) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_solve_two_gpus(tmp_path, exchange):
    """Full C1 solve sharded over 2 GPUs: the per-pivot exchange through NVLink peer mailboxes (default)
    and through ncclAllGather (GNE_EXCHANGE=nccl) must both reproduce the reference's pivot trace."""
    import genneteq_b200 as gne
    if gne.cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = {"GNE_EXCHANGE": "nccl"} if exchange == "nccl" else {}
    port = 29612 if exchange == "p2p" else 29613
    assert launch("nccl", 2, str(tmp_path / (exchange + ".txt")), port, env) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("rule", ["LV", "QS"])
def test_sharded_c2_prefix_window_all_gpus(tmp_path, rule):
    """configs[1] size (100k nodes / 2M arcs) sharded over every GPU of the box (2, 4 or 8): the first 300 pivots,
    flows and potentials equal the oracle's, and every rank holds the same trace."""
    import genneteq_b200 as gne
    ngpu = gne.cuda_device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    assert launch("nccl-c2", nproc, str(tmp_path / ("c2" + rule + ".txt")), 29640 + (1 if rule == "QS" else 0), None, rule) == "OK"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nccl", "nccl-c2"])
def test_sharded_streamed_pass_with_fused_exchange(tmp_path, mode):
    """The TMA-streamed pass whose last CTA runs the reduction AND the NVLink exchange (the path configs[2] takes
    on 2-8 GPUs), forced onto small shards with GNE_LV_KERNEL=stream: the full C1 solve (long tie lists, overflow
    fallbacks, both phases) on 2 GPUs and the configs[1] prefix window on every GPU of the box must reproduce the
    reference / oracle trace on every rank."""
    import genneteq_b200 as gne
    ngpu = gne.cuda_device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 2 if mode == "nccl" else (8 if ngpu >= 8 else (4 if ngpu >= 4 else 2))
    port = 29660 + (1 if mode == "nccl-c2" else 0)
    assert launch(mode, nproc, str(tmp_path / ("fused" + mode + ".txt")), port, {"GNE_LV_KERNEL": "stream"}, "LV") == "OK"
