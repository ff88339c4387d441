"""GPU (needs >= 2 devices; skipped otherwise): the row-partitioned eigensolve over NCCL equals the single-GPU one."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


TRANSPORTS = {"peer": {}, "nccl_halo": {"BH_DIST_PEER": "0"}, "allgather": {"BH_DIST_ALLGATHER": "1"}}


def run_check(args, env_extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "eigs_mgpu.py"), "--check"] + args
    env = dict(os.environ)
    env.update(env_extra)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    return json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("args", [["-m", "10", "-n", "10", "-U", "4", "--nev", "20", "--point"],
                                  ["-m", "12", "-n", "12", "-U", "1", "--nev", "2", "--ncv", "12"]])
def test_partitioned_equals_single(args):
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = run_check(args, {}, 29547)
    assert out["check"] == "ok" and out["world"] == 2


@pytest.mark.parametrize("transport", ["nccl_halo", "allgather"])
def test_fallback_transports_equal_single(transport):
    # the default is the peer-memory form (CUDA IPC arena + copy-engine pulls); the two NCCL forms stay as fall-backs
    # (no peer access / BH_DIST_PEER=0, and the plain all-gather of x) and must give the same eigenpairs
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = run_check(["-m", "12", "-n", "12", "-U", "1", "--nev", "2", "--ncv", "12"], TRANSPORTS[transport], 29549)
    assert out["check"] == "ok" and out["world"] == 2
