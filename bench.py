#!/usr/bin/env python
"""bench.py — particles/sec per SMC set (PLS ranking + top-N selection + doubled variance + weight update).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4|C5] [--impl ours|reference]

A step is one pass of the hot path over one synthetic SMC set of the workload's shape (abcsmc_b200/synth.py,
SURVEY.md §8d): rank all N particles with the PLS filter, keep the top N_pp, gather their parameters, compute the
doubled variance, and update the importance weights against the previous set's predictive prior.
  value : whole-job particles/s with inputs resident in HBM, timed with CUDA events on the launching stream.
  e2e   : the same through the reference-facing host API (abcsmc_b200.api: host buffers in pinned memory, H2D of the
          set and D2H of order / variance / weights inside the timed region).
N > 1 (torchrun, one process per GPU): rank 0 ranks the set; the weight update is row-sharded over all ranks with the
previous set broadcast and the sum of squares all-reduced over NCCL (north_star: only the weight update shards).
--impl reference times the CPU oracle (oracle/abc_oracle.cpp, a restatement: the reference itself cannot be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particles/sec per SMC set (PLS+select+reweight)"
UNIT = "particles/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak_tflops():
    """FP64 DMMA peak measured on this pool's B200 by tools/fp64_peak.cu (profiles/r01_fp64_peak.json)."""
    p = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dmma_tflops"]), "measured (profiles/r01_fp64_peak.json, DMMA.8x8x4 loop)"
    except Exception:
        return 37.0, "nominal 148 SM x 64 lanes x 2 x 1.965 GHz"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step(cfg, orc):
    """One full step of the hot path on the CPU oracle, the reference's call sequence (AbcSmc.cpp:634-664, 1041-1066)."""
    r = orc.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    sel = cfg["params"][r["order"][:cfg["N_pp"]].astype(np.int64), :]
    dv = orc.calculate_doubled_variance(sel)
    w = orc.weight_predictive_prior(np.ones(len(sel)), sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
    return r["order"][:cfg["N_pp"]], dv, w


def oracle_weights_sample(cfg, orc, n_new, n_old):
    orc.weight_predictive_prior(np.ones(n_new), cfg["theta_new"][:n_new], cfg["theta_old"][:n_old], cfg["w_old"][:n_old], cfg["dv_old"])


def make_workload(name):
    from abcsmc_b200 import synth
    if name == "C4":
        c = synth.CONFIGS["C4"]
        th_new, th_old, w_old, dv_old = synth.make_weight_case(c["N_new"], c["N_old"], c["P"], c["seed"])
        return dict(name="C4", N=c["N_new"], P=c["P"], K=0, N_pp=c["N_new"], theta_new=th_new, theta_old=th_old, w_old=w_old, dv_old=dv_old)
    return synth.make_config(name)


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle as orc
    orc.build()
    cfg = make_workload(args.workload)
    cores = 1
    if cfg["name"] == "C4":
        n = 4000   # bounded sample: n x n pairs of the 1M x 1M job, extrapolated by the exact law N_new*N_old*P
        fn = lambda: oracle_weights_sample(cfg, orc, n, n)
        scale = (cfg["N"] / n) ** 2
        sample = f"{n}x{n} pairs of the 1Mx1M weight update, time scaled by (1e6/{n})^2 (law: N_new*N_old*P)"
    else:
        fn = lambda: oracle_step(cfg, orc)
        scale = 1.0
        sample = "full workload per step (one complete set)"
    for _ in range(args.warmup if args.warmup < 2 else 1):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps * scale
    value = cfg["N"] / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(cfg, world):
    c = {"workload": f"{cfg['name']}: N={cfg['N']} particles, P={cfg['P']} params, K={cfg['K']} metrics, top-N={cfg['N_pp']}, "
                     f"pls_training_fraction=0.5, previous predictive prior {cfg['theta_old'].shape[0]} particles",
         "parallelism": "rank 0 ranks the set; weight-update rows sharded over %d GPU(s)" % world,
         "l2": "L2 flushed between steps by writing a 512 MiB buffer (outside the timed region)"}
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", type=int, default=0, help="0 KERNEL_TYPE1 (reference default), 1 KERNEL_TYPE2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 3 and args.workload != "C2":
            args.steps = 3
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from abcsmc_b200 import api, device as dev
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    devt = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=devt)
    ctx = api.Context(local_rank)
    dev.use_torch_stream(ctx)
    W = max(args.warmup, 3)

    cfg = make_workload(args.workload)
    N, P, K, N_pp = cfg["N"], cfg["P"], cfg["K"], cfg["N_pp"]
    is_c4 = cfg["name"] == "C4"

    def pinned(a):   # numpy (rows, cols) -> pinned column-major host buffer, returned as a Fortran-ordered numpy view
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)).pin_memory()
        return t, t.numpy().T

    # host (pinned) and device-resident copies of the step's inputs
    keep = []
    if not is_c4:
        if rank == 0:
            t_met, h_met = pinned(cfg["metrics"]); t_par, h_par = pinned(cfg["params"]); keep += [t_met, t_par]
            d_met = t_met.to(devt); d_par = t_par.to(devt)
            d_target = torch.from_numpy(cfg["target"]).to(devt)
        h_target = cfg["target"]
    else:
        t_new, h_new = pinned(cfg["theta_new"]); keep.append(t_new)
    t_old, h_old = pinned(cfg["theta_old"]); keep.append(t_old)
    d_old = t_old.to(devt)
    d_wold = torch.from_numpy(cfg["w_old"]).to(devt)
    d_dvold = torch.from_numpy(cfg["dv_old"]).to(devt)
    if is_c4:
        d_new = t_new.to(devt)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=devt)

    def step_device():
        """inputs resident in HBM; outputs stay in HBM"""
        if is_c4:
            if world > 1:
                return dev.weights_sharded(ctx, None, d_new, d_old, d_wold, d_dvold, gather=False)
            return dev.weights(ctx, None, d_new, d_old, d_wold, d_dvold)
        if world == 1:
            order, _, used, _ = dev.rank_pls(ctx, d_met, d_par, d_target, 0.5, top_n=N_pp, method=args.method)
            g, dv = dev.doubled_variance_gather(ctx, d_par, order)
            return dev.weights(ctx, None, g, d_old, d_wold, d_dvold)
        g = torch.empty((P, N_pp), dtype=torch.float64, device=devt)
        if rank == 0:
            order, _, used, _ = dev.rank_pls(ctx, d_met, d_par, d_target, 0.5, top_n=N_pp, method=args.method)
            g0, dv = dev.doubled_variance_gather(ctx, d_par, order)
            g.copy_(g0)
        dist.broadcast(g, 0)                      # the new predictive prior's parameters (N_pp x P)
        dist.broadcast(d_old, 0); dist.broadcast(d_wold, 0); dist.broadcast(d_dvold, 0)   # previous set, as north_star states
        return dev.weights_sharded(ctx, None, g, d_old, d_wold, d_dvold)

    h2d_bytes = d2h_bytes = 0

    def step_host():
        """the reference-facing call sequence on host buffers (AbcSmc.cpp:634-664, 1041-1066)"""
        nonlocal h2d_bytes, d2h_bytes
        if is_c4:
            if world > 1:
                dn = t_new.to(devt, non_blocking=True); do = t_old.to(devt, non_blocking=True)
                w = dev.weights_sharded(ctx, None, dn, do, d_wold, d_dvold, gather=False)
                h2d_bytes = (t_new.numel() + t_old.numel()) * 8; d2h_bytes = w.numel() * 8
                return w.cpu()
            h2d_bytes = (t_new.numel() + t_old.numel() + N + P) * 8; d2h_bytes = N * 8
            return api.weight_predictive_prior(None, h_new, h_old, cfg["w_old"], cfg["dv_old"], ctx=ctx)
        if world == 1:
            order = api.particle_ranking_PLS(h_met, h_par, h_target, 0.5, top_n=N_pp, method=args.method, ctx=ctx)
            sel = np.asfortranarray(h_par[order.astype(np.int64), :])
            dv = api.calculate_doubled_variance(sel, ctx=ctx)
            w = api.weight_predictive_prior(None, sel, h_old, cfg["w_old"], cfg["dv_old"], ctx=ctx)
            h2d_bytes = (N * (K + P) + K + 2 * N_pp * P + h_old.size + cfg["w_old"].size + P) * 8
            d2h_bytes = (N_pp + P + N_pp) * 8
            return w
        g = torch.empty((P, N_pp), dtype=torch.float64, device=devt)
        if rank == 0:
            order = api.particle_ranking_PLS(h_met, h_par, h_target, 0.5, top_n=N_pp, method=args.method, ctx=ctx)
            sel = np.asfortranarray(h_par[order.astype(np.int64), :])
            dv = api.calculate_doubled_variance(sel, ctx=ctx)
            g.copy_(torch.from_numpy(np.ascontiguousarray(sel.T)))
        do = t_old.to(devt, non_blocking=True)
        dist.broadcast(g, 0); dist.broadcast(do, 0); dist.broadcast(d_wold, 0); dist.broadcast(d_dvold, 0)
        w = dev.weights_sharded(ctx, None, g, do, d_wold, d_dvold)
        h2d_bytes = (N * (K + P) + K + 2 * N_pp * P + h_old.size) * 8; d2h_bytes = (N_pp + P + N_pp) * 8
        return w.cpu() if rank == 0 else None

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        total_ms = 0.0
        stage_acc = {}
        n_launch0 = ctx.launches
        for _ in range(steps):
            flush.fill_(1)                                # L2 flush, outside the timed region
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            for k, v in ctx.stage_ms().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
        t = torch.tensor([total_ms], dtype=torch.float64, device=devt)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, {k: v / steps for k, v in stage_acc.items()}, ctx.launches - n_launch0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, stages, launches = timed(step_device, args.steps, W)
    ms_e2e, stages_e2e, launches_e2e = timed(step_host, args.steps, W)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        hbm_peak, hbm_src = load_peaks()
        # dominant stage -> roofline of its dominant kernel (formulas: SURVEY.md §8d, restated in DESIGN.md)
        dom = max((k for k in stages if k not in ("h2d", "d2h")), key=lambda k: stages[k])
        N_old = cfg["theta_old"].shape[0]
        if dom == "weight_update":
            shard = (N_pp + world - 1) // world
            flops = shard * N_old * (3 * P + 2)
            peak, src = fp64_peak_tflops()
            ach = flops / (stages[dom] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "weights_dmma_kernel (FP64 DMMA.8x8x4 + FP64 exp epilogue)", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                    "note": f"algorithmic (3P+2) flop per pair x {shard}x{N_old} pairs per launch, exp counted separately; stage time incl. pack kernels; FP64 peak {src}"}
        else:
            n_tr = int(round(N * 0.5)); n_te = N - n_tr
            alg = {"moments_zscore": 3 * 8 * N * (K + P), "pls_fit": 8 * n_tr * (K + P) + K * 8 * n_tr * (K + 1),
                   "holdout_press": 8 * n_te * (K + P), "wilcoxon_select": None, "project_distance": 8 * N * K + 8 * N,
                   "ordering": 8 * N, "doubled_variance": 8 * N_pp * P}[dom]
            if alg is None:
                roof = {"bound": "hbm", "kernel": "radix_scatter_kernel (batched Wilcoxon sorts)", "achieved": None, "peak": hbm_peak, "unit": "GB/s",
                        "frac": None, "traffic": None, "note": "test count is data dependent; see profiles/ for the per-pass GB/s"}
            else:
                ach = alg / (stages[dom] * 1e-3) / 1e9
                roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                        "note": f"stage-level: algorithmic bytes of SURVEY §8d over the stage's CUDA-event time; HBM peak {hbm_src}"}
        line = {"metric": METRIC, "value": N / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, world), "clocks": clocks,
                "e2e": {"value": N / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes)},
                "gpu_launches": int(launches), "gpu_launches_e2e": int(launches_e2e), "stages_ms": stages, "stages_ms_e2e": stages_e2e, "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            import oracle as orc
            orc.build()
            if is_c4:
                n = 3000
                t0 = time.perf_counter(); oracle_weights_sample(cfg, orc, n, n); dt = (time.perf_counter() - t0) * (N / n) ** 2
                sample = f"{n}x{n} pairs, scaled by (N/{n})^2 (law N_new*N_old*P)"
            else:
                reps, t0 = 0, time.perf_counter()
                while reps < 3 and (time.perf_counter() - t0 < 10.0 or reps == 0):
                    oracle_step(cfg, orc); reps += 1
                dt = (time.perf_counter() - t0) / reps
                sample = f"{reps} full step(s) of the same workload"
            line["cpu_baseline"] = {"value": N / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                                    "host_cores_available": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
