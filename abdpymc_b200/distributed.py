"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` for plumbing.

Two axes shard naturally:

* **individuals** -- given the 17 scalars every likelihood term, Bernoulli term and Gibbs update
  of individual n depends only on column n (abd.py:343, 393 gather per (gap, ind); the
  constraints abd.py:560-862 are column-wise).  Rank r holds a contiguous block of individuals
  (``CohortArrays.shard``) and the matching slice of every chain's ``i_raw`` / ``waner``.  One
  evaluation = local raw sums (``abd_sums_dev``) -> ONE all-reduce of C x 16 doubles ->
  ``abd_finalize_logp_dev`` on every rank.  The Gibbs sweep needs no collective at all.
* **chains** -- fully independent: replicate the cohort, give each rank its own chains, no
  collective on the data path.

The reference has no distributed code of any kind (pm.sample(cores=K) forks independent chain
processes, abd.py:922); this module is the B200-native counterpart of that chain-level process
parallelism plus the individual sharding large cohorts need.
"""

from __future__ import annotations

import os

import numpy as np

from .cohort import CohortArrays, shard_bounds

N_SUMS = 16


def dist_env():
    """(rank, world, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None, device=None):
    """Initialise torch.distributed from the torchrun environment; no-op for a single process.
    NCCL when a CUDA device is given, gloo otherwise (CPU tests)."""
    import torch.distributed as dist

    rank, world, _ = dist_env()
    if world == 1 or dist.is_initialized():
        return dist if world > 1 else None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    backend = backend or ("nccl" if device is not None else "gloo")
    kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return dist


def chain_slice(n_chains: int, rank: int, world: int) -> slice:
    """Chains owned by ``rank`` when chains are sharded (contiguous, sizes differ by <= 1)."""
    lo, hi = shard_bounds(n_chains, rank, world)
    return slice(lo, hi)


def cohort_totals(cohort: CohortArrays) -> tuple:
    """(n_inds, n_rows_s, n_rows_n) of the whole cohort: what every shard's engine must be told
    so that the finaliser uses global Bernoulli / likelihood constants."""
    return cohort.n_inds, int((cohort.antigen == 1).sum()), int((cohort.antigen == 0).sum())


def shard_cohort(cohort: CohortArrays, rank: int, world: int):
    """(shard, totals, ind_offset, slice of individuals) for individual sharding."""
    lo, hi = shard_bounds(cohort.n_inds, rank, world)
    return cohort.shard(rank, world), cohort_totals(cohort), lo, slice(lo, hi)


def allreduce_sums(sums, group=None):
    """Sum the per-shard raw sums (C x 16 doubles, every entry additive over individuals) over
    all ranks, in place.  ``sums``: torch tensor on the device the process group runs on."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


class ShardedEngine:
    """Individual-sharded evaluation of the joint logp + gradient on this rank's GPU.

    Every rank builds the engine of its shard with the global totals; ``logp_dlogp`` runs
    sums -> all-reduce -> finalise and leaves identical (logp, dlogp) tensors on every rank.
    """

    def __init__(self, cohort: CohortArrays, splits=None, ignore_pcrpos=False, device_index=0, rank=None, world=None,
                 group=None, fused=False, max_chains=64):
        """``fused=True``: the all-reduce is done INSIDE the logp kernel over NVLink peer memory
        (``abd_logp_dlogp_sharded_dev``: every rank's finishing CTA stores its sums into all peers'
        exchange buffers and waits for theirs) -- one launch per evaluation, no NCCL call.  Needs
        an initialised process group once, to exchange the CUDA IPC handles."""
        import torch

        from .engine import AbdEngine

        env_rank, env_world, _ = dist_env()
        self.rank = env_rank if rank is None else rank
        self.world = env_world if world is None else world
        self.group = group
        shard, totals, offset, self.ind_slice = shard_cohort(cohort, self.rank, self.world)
        self.device = torch.device("cuda", device_index)
        self.engine = AbdEngine(shard, splits=splits, ignore_pcrpos=ignore_pcrpos, device=device_index, totals=totals,
                                ind_offset=offset)
        self.G, self.N_local, self.N_total = self.engine.G, self.engine.N, cohort.n_inds
        self._bufs = {}
        self.fused = bool(fused) and self.world > 1
        if self.fused:
            import torch.distributed as dist

            mine = self.engine.xch_alloc(self.world, self.rank, max_chains)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            self.engine.xch_connect(handles)
            dist.barrier(group=group)  # every rank has mapped every buffer before the first launch

    def _buffers(self, C):
        import torch

        if C not in self._bufs:
            self._bufs[C] = (
                torch.zeros(C, N_SUMS, dtype=torch.float64, device=self.device),
                torch.zeros(C, dtype=torch.float64, device=self.device),
                torch.zeros(C, 17, dtype=torch.float64, device=self.device),
            )
        return self._bufs[C]

    def upload_state(self, i_raw, waner):
        """``i_raw`` (C, G, N_total), ``waner`` (C, N_total): keeps this rank's individuals."""
        i_raw, waner = np.asarray(i_raw), np.asarray(waner)
        return self.engine.upload_state(np.ascontiguousarray(i_raw[..., self.ind_slice]),
                                        np.ascontiguousarray(waner[..., self.ind_slice]))

    def logp_dlogp(self, q17, n_chains=None):
        """``q17``: torch (C, 17) float64 tensor on this rank's device (identical on all ranks).
        Uses the resident chain state.  Returns (logp (C,), dlogp (C, 17)) device tensors."""
        import torch

        C = q17.shape[0] if n_chains is None else n_chains
        sums, out, outg = self._buffers(C)
        st = torch.cuda.current_stream(self.device).cuda_stream
        d_i, d_w = self.engine.state_dev(C)
        if self.fused:
            self.engine.logp_dlogp_sharded_dev(C, q17.data_ptr(), d_i, d_w, out.data_ptr(), outg.data_ptr(), st)
            return out, outg
        self.engine.sums_dev(C, q17.data_ptr(), 1, d_i, d_w, sums.data_ptr(), st)
        allreduce_sums(sums, self.group)
        self.engine.finalize_logp_dev(C, q17.data_ptr(), sums.data_ptr(), out.data_ptr(), outg.data_ptr(), st)
        return out, outg

    def gibbs_sweep(self, q17, seed, sweep, mode=0, transit_p=0.8):
        """Sweep this rank's individuals of every chain in place (no collective: theta is the
        same on every rank, updates are local; RNG streams are keyed by GLOBAL individual)."""
        import torch

        C = q17.shape[0]
        st = torch.cuda.current_stream(self.device).cuda_stream
        d_i, d_w = self.engine.state_dev(C)
        self.engine.gibbs_sweep_dev(C, q17.data_ptr(), 1, None, None, d_i, d_w, seed, sweep, mode=mode,
                                    transit_p=transit_p, stream=st)

    def close(self):
        self.engine.close()
