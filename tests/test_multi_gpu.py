"""On-hardware multi-GPU correctness: one Baum-Welch iteration on 2 GPUs (torchrun, NCCL) gives the compiled
reference's updated parameters, and the statistics of the sharded run -- ONE all_gather of one packed device buffer --
agree with a single GPU's within 1e-12 sqrt(S).  Skipped on boxes with fewer than two GPUs."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_baum_welch_on_two_gpus_matches_one_gpu_and_the_reference():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_bw_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("DIST_BW ")][-1]
    out = json.loads(line[len("DIST_BW "):])
    assert out["world"] == 2 and out["events"] == 6000
    assert out["golden_dense_ok"] and out["golden_params_ok"]
    assert out["edge_rel"] <= out["tol"] and out["moment_rel"] <= out["tol"] and out["loglik_rel"] <= out["tol"]
    assert out["minmax_equal"] and out["ranks_identical"]
