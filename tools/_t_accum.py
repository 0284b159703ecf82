import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
from abdpymc_b200.abd import infer_builtin
from abdpymc_b200.cohort import synthetic_cohort
co = synthetic_cohort(10000)
res, post, last = infer_builtin(co, (14, 20), False, tune=300, draws=700, chains=4, seed=1)
print("infer_builtin 10k, deterministics recorded every draw:", (300+700)/res.wall_s, "it/s; mean_i sum", res.means["i"].sum(), "true", co.truth["infections"].sum())
