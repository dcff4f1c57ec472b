#!/bin/bash
TAG=${1:-r2h}; O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu_$TAG.log | head -30 | cut -c1-300
python - <<'P' 2>&1 | tee $O/h2d_stage_$TAG.txt
import numpy as np, time, torch
from abcsmc_b200 import api, synth
ctx = api.get_context(0)
for name in ("C3", "T1M"):
    cfg = synth.make_config(name)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory()
        return t, t.numpy().T
    tm, hm = pin(cfg["metrics"]); tp, hp = pin(cfg["params"])
    ctx.set_timers(stages=True, kernels=())
    for i in range(4):
        t0 = time.perf_counter()
        api.particle_ranking_PLS(hm, hp, cfg["target"], 0.5, top_n=cfg["N_pp"], ctx=ctx)
        dt = (time.perf_counter() - t0) * 1e3
        st = ctx.stage_ms()
        print(name, f"call {dt:.2f} ms; h2d {st.get('h2d', 0):.2f} ms = {(hm.nbytes + hp.nbytes) / st['h2d'] / 1e6:.1f} GB/s; stages", {k: round(v, 2) for k, v in st.items() if v})
P
bash tools/gpu_ab.sh $TAG "C3 T1M C2" "base"
