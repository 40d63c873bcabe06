"""GPU, >= 2 devices: real ranks, real NCCL (tests/multirank_worker.py under torchrun).  Skipped on a one-GPU box;
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu` runs it."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_over_nccl_match_one_rank():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert r.stdout.count("multirank checks OK") == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_alds_routed_predict_matches_one_rank():
    """BASELINE configs 3 / 5: PCA + k-means routing to 4 per-cluster models, the subdomains shared cluster-major by two
    real ranks (one in-place all-gather on libfesr's communicator): stitched field, subdomain weights and labels equal
    the single-rank result bit for bit (tools/check_alds_multi.py is the body of the check)."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_alds_multi.py"),
           "--mesh-n", "20", "--steps", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{"check"')]
    assert line, r.stdout[-2000:]
    res = json.loads(line[-1])
    assert res["ok"] and res["n_gpus"] == 2
    assert res["field_rel_l2_vs_single_rank"] == 0.0 and res["weight_rel_err"] == 0.0
    assert min(res["labels_hist"]) > 0          # every cluster is in use
