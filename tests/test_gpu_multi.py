"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): individual sharding with the
all-reduce fused into the logp kernel over NVLink peer memory against the NCCL path and the
unsharded engine.  The single-process logic of the sharding is covered on CPU (gloo) in
test_host_logic.py."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_peer_allreduce_matches_nccl_and_unsharded():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29593", str(ROOT / "tools" / "xch_check.py"), "6000", "3"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and line["rel_err_vs_unsharded"] < 1e-11
