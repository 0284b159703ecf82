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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_peer_allreduce_with_multi_chain_ctas_and_compact_cells():
    """Shards of 25 000 individuals x 4 chains: the planner runs them as one wave of two-chain CTAs with the compact
    cell layout, i.e. the exchange is posted from the deferred (after the chain loop) finaliser."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29595", str(ROOT / "tools" / "xch_check.py"), "50000", "4"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and line["rel_err_vs_unsharded"] < 1e-11
    assert line["plan_rank0"]["compact_cells"] == 1 and line["plan_rank0"]["chains_per_cta"] == 2, line


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_compound_sampler_on_an_individual_sharded_cohort():
    """HMC + Gibbs with the individuals split over 2 GPUs (distributed.ShardedTarget: fused NVLink all-reduce
    inside every leapfrog launch, local Gibbs sweeps) against the same sampler on one GPU."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29597", str(ROOT / "tools" / "sharded_sampler_check.py"), "3000", "2", "30", "30"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and line["fused_identical_on_all_ranks"] and line["nccl_identical_on_all_ranks"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cli_over_two_devices(tmp_path):
    """abdpymc-infer --devices 0,1 with both shardings writes the reference's variable names."""
    import numpy as np

    from abdpymc_b200.cohort import CohortArrays

    CohortArrays.load("cohort").to_disk(tmp_path / "cohort_data")
    for shard in ("individuals", "chains"):
        out = tmp_path / f"post_{shard}.npz"
        cmd = [sys.executable, "-m", "abdpymc_b200.abd", "--tune", "30", "--draws", "20", "--ititers_data",
               str(tmp_path / "cohort_data"), "--split_delta", "--split_omicron", "--netcdf", str(out), "--chains", "4",
               "--devices", "0,1", "--shard", shard, "--thinned", "0"]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        z = np.load(out)
        assert z["ab_n_perm"].shape == (4, 20) and z["last_i_raw"].shape == (4, 31, 1520) and z["mean_i"].shape == (31, 1520)
        assert np.isfinite(z["sample_stats_lp"]).all()
