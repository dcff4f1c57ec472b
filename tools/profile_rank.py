"""Runs the ranking path twice (warm-up + measured) on a BASELINE.json workload; meant to be wrapped by ncu:
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
       python tools/profile_rank.py C3
(the warm-up repetitions are not captured: cudaProfilerStart is called before the last one)
Prints the number of kernels each call launched so that -s/-c can be set."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abcsmc_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = synth.make_config(name)
ctx = api.get_context(0)
ctx.set_timers(stages=True, kernels="all")
import ctypes
_rt = ctypes.CDLL("libcudart.so")          # ncu --profile-from-start off: only the last repetition is captured
for i in range(reps):
    if i == reps - 1:
        _rt.cudaProfilerStart()
    l0 = ctx.launches
    order = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=cfg["N_pp"], ctx=ctx)
    sel = np.asfortranarray(cfg["params"][order.astype(np.int64), :])
    dv = api.calculate_doubled_variance(sel, ctx=ctx)
    w = api.weight_predictive_prior(None, sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"], ctx=ctx)
    print(f"call {i}: {ctx.launches - l0} launches; stages {ctx.stage_ms()}")
