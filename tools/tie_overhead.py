"""What abcb200_set_tie_order(1) costs when there is nothing to re-derive (continuous metrics): abcb200_rank_pls on pinned host buffers at a
BASELINE.json shape, mode 0 vs mode 1, wall clock around the synchronous call. usage: python tools/tie_overhead.py [C3] [reps]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from abcsmc_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = synth.make_config(name)


def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory()
    return t, t.numpy().T


t_met, h_met = pinned(cfg["metrics"]); t_par, h_par = pinned(cfg["params"])
ctx = api.get_context(0)
out = {}
for mode in (0, 1, 0, 1):
    ctx.set_tie_order(mode)
    for _ in range(3):
        o = api.particle_ranking_PLS(h_met, h_par, cfg["target"], 0.5, top_n=cfg["N_pp"], ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(reps):
        o = api.particle_ranking_PLS(h_met, h_par, cfg["target"], 0.5, top_n=cfg["N_pp"], ctx=ctx)
    out.setdefault(mode, []).append((time.perf_counter() - t0) / reps * 1e3)
    out[("order", mode)] = o
ctx.set_tie_order(0)
assert np.array_equal(out[("order", 0)], out[("order", 1)])
print(f"{name}: abcb200_rank_pls end to end, ms per call: ties by index {min(out[0]):.3f}, std::sort placement {min(out[1]):.3f} "
      f"(re-derived orders: {ctx.tie_resorts}); N={cfg['N']} top_n={cfg['N_pp']}")
