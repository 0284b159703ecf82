#!/usr/bin/env python3
"""Timing sweeps of the kernels on one GPU (developer tool; not part of the bench contract).
usage: python tools/tune.py [sums|gibbs|floor] ..."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402


def graph_time(fn, n_inner=128, reps=10):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fn(side.cuda_stream)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(n_inner):
            fn(side.cuda_stream)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n_inner) * 1e3  # us per call


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "sums"
    n_inds = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    chains = [int(c) for c in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["4"])]
    dev = torch.device("cuda:0")
    for C in chains:
        co, q, vals, i_raw, w = bench.workload(n_inds=n_inds, n_chains=C)
        eng = AbdEngine(co, splits=bench.SPLITS)
        eng.upload_state(i_raw, w)
        di, dw = eng.state_dev(C)
        tq = torch.from_numpy(q).to(dev)
        out = torch.zeros(C, dtype=torch.float64, device=dev)
        outg = torch.zeros(C, 17, dtype=torch.float64, device=dev)
        sums = torch.zeros(C, 16, dtype=torch.float64, device=dev)
        if what == "floor":
            t = graph_time(lambda st: eng.finalize_logp_dev(C, tq.data_ptr(), sums.data_ptr(), out.data_ptr(), outg.data_ptr(), st))
            print(f"C={C} k_finalize (launch floor) {t:.2f} us")
        elif what == "sums":
            for rows in [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["0"])]:
                for cpc in [int(v) for v in (sys.argv[5].split(",") if len(sys.argv) > 5 else ["0"])]:
                    eng.set_tuning(rows, cpc)
                    t = graph_time(lambda st: eng.logp_dlogp_dev(C, tq.data_ptr(), di, dw, out.data_ptr(), outg.data_ptr(), st),
                                   n_inner=64 if C <= 16 else 4)
                    print(f"N={n_inds} C={C} rows/tile={rows} chains/cta={cpc}: {t:.2f} us/launch, {C / t * 1e6:.0f} evals/s, "
                          f"{eng.algorithmic_bytes_logp(C) / t / 1e3:.1f} GB/s")
        elif what == "traj":
            a = np.random.default_rng(0).normal(size=(17, 17))
            tm = torch.from_numpy(1e-6 * (a @ a.T / 17 + np.eye(17))).to(dev)
            te = torch.full((C,), 1e-3, dtype=torch.float64, device=dev)
            eng.logp_dlogp_dev(C, tq.data_ptr(), di, dw, out.data_ptr(), outg.data_ptr(), 0)
            torch.cuda.synchronize()
            g0 = outg.clone()
            for L in [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["1", "16", "64", "256"])]:
                qq, pp, gg = tq.clone(), torch.zeros_like(tq), g0.clone()
                lp = torch.zeros(C, dtype=torch.float64, device=dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ts = []
                for rep in range(6):
                    qq.copy_(tq), pp.zero_(), gg.copy_(g0)
                    e0.record()
                    eng.leapfrog_dev(C, L, qq.data_ptr(), pp.data_ptr(), gg.data_ptr(), lp.data_ptr(), te.data_ptr(), tm.data_ptr(), di, dw, 0)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                eng.leapfrog_status(C)
                t = float(np.median(ts[1:]))
                print(f"N={n_inds} C={C} persistent trajectory L={L}: {t:.1f} us/launch, {t / L:.2f} us/step, {C * L / t * 1e6:.0f} evals/s, logp {lp[0].item():.3f}")
        elif what == "gibbs":
            th = torch.from_numpy(vals[:, [1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 14, 15, 16]].copy()).to(dev)
            p = torch.from_numpy(vals[:, 0].copy()).to(dev)
            pw = torch.from_numpy(vals[:, 7].copy()).to(dev)
            for mode in (0, 1):
                k = [0]

                def fn(st):
                    k[0] += 1
                    eng.gibbs_sweep_dev(C, th.data_ptr(), 0, p.data_ptr(), pw.data_ptr(), di, dw, 1, k[0], mode=mode, stream=st)

                t = graph_time(fn, n_inner=8, reps=5)
                print(f"N={n_inds} C={C} mode={mode}: {t:.1f} us/sweep-launch, {C / t * 1e6:.0f} sweeps/s")
        eng.close()


if __name__ == "__main__":
    main()
