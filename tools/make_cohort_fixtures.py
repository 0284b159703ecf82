#!/usr/bin/env python3
"""Pack the reference's two bundled cohorts into compact .npz fixtures.

Run ONCE in the build container (needs /root/reference, which does not exist on the GPU box):

    python tools/make_cohort_fixtures.py

Reads   /root/reference/data/{cohort_data,test_data/cohort_data}/{df.csv,vacs.txt,pcrpos.txt,t0.txt}
exactly the way ``TiterData.from_disk`` does (abd.py:171-202) and keeps only the columns the
inference hot path consumes (abd.py:35-36, 82-98, 413-418, 462, 468):
individual_i, elapsed_months, antigen (measurement == '10222020-S' -> S, '40588-V08B' -> N),
log_dilution, od; plus vacs / pcrpos (n_inds, n_gaps) and t0.  Row order is df.csv's.

Writes  abdpymc_b200/data/cohort.npz       (1520 individuals x 31 gaps, 35 709 OD rows)
        abdpymc_b200/data/test_cohort.npz  (10 individuals x 26 gaps, 288 OD rows)
These are DATA fixtures (the schedule every synthetic benchmark cohort is bootstrapped from,
SURVEY.md section 8d); no reference source code is copied.
"""
from pathlib import Path

import numpy as np
import pandas as pd

REF = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent.parent / "abdpymc_b200" / "data"

S_NAME, N_NAME = "10222020-S", "40588-V08B"


def pack(src: Path, dst: Path) -> None:
    df = pd.read_csv(src / "df.csv", index_col=0)
    vacs = np.loadtxt(src / "vacs.txt")
    pcrpos = np.loadtxt(src / "pcrpos.txt")
    t0 = (src / "t0.txt").read_text().strip()
    assert set(df["measurement"].unique()) <= {S_NAME, N_NAME}
    assert set(np.unique(vacs)) <= {0.0, 1.0} and set(np.unique(pcrpos)) <= {0.0, 1.0}
    assert vacs.shape == pcrpos.shape
    np.savez_compressed(
        dst,
        t0=np.array(t0),
        vacs=vacs.astype(np.uint8),
        pcrpos=pcrpos.astype(np.uint8),
        ind=df["individual_i"].to_numpy(np.int32),
        gap=df["elapsed_months"].to_numpy(np.int32),
        antigen=(df["measurement"] == S_NAME).to_numpy(np.uint8),  # 0 = N, 1 = S
        x=df["log_dilution"].to_numpy(np.float64),
        od=df["od"].to_numpy(np.float64),
        record_id=df["record_id"].to_numpy(np.int64),
    )
    print(dst, vacs.shape, len(df), f"{dst.stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    pack(REF / "cohort_data", OUT / "cohort.npz")
    pack(REF / "test_data" / "cohort_data", OUT / "test_cohort.npz")
