#!/usr/bin/env python3
"""Individual sharding over >= 2 GPUs of one node: the NCCL path (abd_sums_dev -> all_reduce ->
abd_finalize_logp_dev) against the fused peer-memory path (abd_logp_dlogp_sharded_dev), both
against the unsharded engine, with timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29591 tools/xch_check.py [n_inds] [n_chains]

Prints one JSON line on rank 0; exits non-zero on any mismatch.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from abdpymc_b200.cohort import synthetic_cohort  # noqa: E402
from abdpymc_b200.distributed import ShardedEngine  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402
import bench  # noqa: E402

n_inds = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

co = synthetic_cohort(n_inds)
rng = np.random.default_rng(5)
i_raw = (rng.random((C, co.n_gaps, co.n_inds)) < 0.04).astype(np.int8)
w = (rng.random((C, co.n_inds)) < 0.5).astype(np.int8)
q = torch.from_numpy(bench.workload(n_inds=1000, n_chains=C)[1]).to(dev)

se_nccl = ShardedEngine(co, splits=bench.SPLITS, device_index=local, rank=rank, world=world)
se_fused = ShardedEngine(co, splits=bench.SPLITS, device_index=local, rank=rank, world=world, fused=True, max_chains=C)
for se in (se_nccl, se_fused):
    se.upload_state(i_raw, w)

lp_n, g_n = (t.clone() for t in se_nccl.logp_dlogp(q))
res = []
for it in range(5):  # several rounds: both buffer parities, sequence numbers advancing
    lp_f, g_f = (t.clone() for t in se_fused.logp_dlogp(q + 1e-3 * it))
    res.append((lp_f, g_f))
lp_f, g_f = res[0]
torch.cuda.synchronize()
se_fused.engine.xch_status()

ok = True
# (1) every rank holds bitwise the same result
both = torch.cat([lp_f, g_f.reshape(-1)])
gathered = [torch.empty_like(both) for _ in range(world)]
dist.all_gather(gathered, both)
ok &= all(torch.equal(gathered[0], t) for t in gathered)
# (2) fused == NCCL path up to the summation order of `world` partial sums
rel = lambda a, b: float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())  # noqa: E731
e_lp, e_g = rel(lp_f, lp_n), float(((g_f - g_n).abs() / g_n.abs().amax(dim=1, keepdim=True)).max())
ok &= e_lp < 1e-13 and e_g < 1e-12
# (3) against the unsharded engine on rank 0
e_full = None
if rank == 0:
    with AbdEngine(co, splits=bench.SPLITS, device=local) as eng:
        lp_1, g_1 = eng.logp_dlogp(q.cpu().numpy(), i_raw, w)
    e_full = max(float(np.max(np.abs(lp_f.cpu().numpy() - lp_1) / np.abs(lp_1))),
                 float(np.max(np.abs(g_f.cpu().numpy() - g_1) / np.abs(g_1).max(axis=1, keepdims=True))))
    ok &= e_full < 1e-11


def timed(fn, n):
    for _ in range(20):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / n * 1e3  # us per evaluation batch


us_nccl = timed(lambda: se_nccl.logp_dlogp(q), 300)
us_fused = timed(lambda: se_fused.logp_dlogp(q), 300)
# the fused path holds no library call: a block of evaluations is one CUDA graph
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    se_fused.logp_dlogp(q)
side.synchronize()
dist.barrier()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
    for _ in range(64):
        se_fused.logp_dlogp(q)
se_fused.engine.xch_stats(reset=True)
us_graph = timed(graph.replay, 20) / 64
ns_wait, n_x = se_fused.engine.xch_stats(reset=True)
wait_vec = torch.tensor(ns_wait[:world].astype(np.float64) / max(n_x, 1) / 1e3, device=dev)   # us waited for each peer
all_wait = [torch.empty_like(wait_vec) for _ in range(world)]
dist.all_gather(all_wait, wait_vec)
se_fused.engine.xch_status()
lp_g = se_fused.logp_dlogp(q)[0]
torch.cuda.synchronize()
ok &= bool(torch.equal(lp_g, lp_f))

flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(flag.item()), "world": world, "n_inds": n_inds, "chains": C, "rel_err_logp_fused_vs_nccl": e_lp,
                      "rel_err_grad_fused_vs_nccl": e_g, "rel_err_vs_unsharded": e_full, "us_per_eval_nccl": us_nccl,
                      "us_per_eval_fused": us_fused, "us_per_eval_fused_graph": us_graph, "plan_rank0": se_fused.engine.last_plan(),
                      "wait_us_rank_by_peer": [[round(float(v), 2) for v in w_] for w_ in all_wait]}))
for se in (se_nccl, se_fused):
    se.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
