#!/usr/bin/env python
"""bench.py — particles/sec per SMC set (PLS ranking + top-N selection + doubled variance + weight update).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4|C5|T1M] [--impl ours|reference]

Default workload: C3 = BASELINE.json configs[2] (250k particles, 30 params, 150 metrics, top-N 5k), the shape the north star's
30 / 150 target is quoted on and the largest of the named ranking configs whose CPU pass fits the reference arm's budget. At
N = 1 the same line carries `configs`: the other ranking shapes measured the same way with fewer steps — C2 (configs[1]), C5
(configs[4], hold-out selection as AbcSmc does it) and T1M (the north-star target: 1M particles x (30 params, 150 metrics),
top-N 10k) — each with ms_per_step, e2e, stage times and per-kernel roofline entries (--no-extra skips them).

A step is one pass of the hot path over one synthetic SMC set of the workload's shape (abcsmc_b200/synth.py,
SURVEY.md §8d): rank all N particles with the PLS filter, keep the top N_pp, gather their parameters, compute the
doubled variance, and update the importance weights against the previous set's predictive prior.
  value : whole-job particles/s with inputs resident in HBM, timed with CUDA events on the launching stream.
  e2e   : the same through the reference-facing host API (abcsmc_b200.api: host buffers in pinned memory, H2D of the
          set and D2H of order / variance / weights inside the timed region).
  roofline : the kernel with the largest share of the step, its CUDA-event time measured live in the timed region,
          against its algorithmic bytes / flops; `roofline_kernels` lists every instrumented kernel.
N > 1 (torchrun, one process per GPU). The ranking stages do not shard ("replicas only"): every rank
processes its own independent SMC set, no collective on the data path, scaling "weak". The one stage that shards, the
weight update, is measured on the C4 stress shape (N_new = N_old = 1M, P = 30) with new-particle rows split over the
ranks, the previous set broadcast and the sum of squares all-reduced over NCCL; it is reported in `sharded_weight_update`
(and is the whole step with --workload C4, scaling "strong").
--impl reference times the CPU oracle (oracle/abc_oracle.cpp, a restatement pinned to the reference's own sources, which only
compile here against Eigen / GSL stand-in headers whose naive inner loops make that build 1.8x SLOWER than the port — DESIGN.md §3)
on the host, single thread like the reference. For C2 and C3 it runs the FULL workload (C3: one pass of
about 200 s whatever --steps says, so the driver's ratio is measured, not extrapolated); C4 / C5 / T1M run a bounded sample.
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


def load_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if "hbm_gbs" in d:
            return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    return 6550.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def load_fp64_peak():
    """FP64 DMMA peak measured on this pool's B200 by tools/fp64_peak.cu (MEASURED_PEAKS.json has no FP64 figure)."""
    p = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dmma_tflops"]), "profiles/r01_fp64_peak.json (tools/fp64_peak.cu, register-resident DMMA.8x8x4 loop)"
    except Exception:
        return 37.0, "nominal 148 SM x 64 FP64 lanes x 2 x 1.965 GHz"


def load_traffic():
    """dram bytes per launch from committed `ncu --set full` captures (profiles/r02_traffic.json), keyed workload -> kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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


# ---- workloads ----------------------------------------------------------------------------------------------------------
def make_workload(name, replica=0):
    from abcsmc_b200 import synth
    if name == "C4":
        c = synth.CONFIGS["C4"]
        th_new, th_old, w_old, dv_old = synth.make_weight_case(c["N_new"], c["N_old"], c["P"], c["seed"])
        return dict(name="C4", N=c["N_new"], P=c["P"], K=0, N_pp=c["N_new"], theta_new=th_new, theta_old=th_old, w_old=w_old, dv_old=dv_old)
    cfg = synth.make_config(name, seed_offset=1000 * replica)
    return cfg


def workload_config(cfg, world, mode):
    if cfg["name"] == "C4":
        wl = f"C4: weight update only, N_new=N_old={cfg['N']}, P={cfg['P']} params"
        par = f"new-particle rows sharded over {world} GPU(s); previous set broadcast, sum of squares all-reduced (NCCL)"
    else:
        wl = (f"{cfg['name']}: N={cfg['N']} particles, P={cfg['P']} params, K={cfg['K']} metrics, top-N={cfg['N_pp']}, "
              f"pls_training_fraction=0.5, previous predictive prior {cfg['theta_old'].shape[0]} particles")
        par = ("one GPU" if world == 1 else
               f"replicas: {world} independent SMC sets, one per GPU, no data-path collective (the ranking does not shard, SURVEY.md 8e)")
    return {"workload": wl, "parallelism": par,
            "l2": "L2 flushed between steps by writing a 512 MiB buffer (outside the timed region)"}


# ---- CPU oracle legs --------------------------------------------------------------------------------------------------------
def oracle_sample(cfg, frac):
    """The rows of `cfg` the CPU leg runs: the first frac*N particles (and frac*N_pp new rows of the weight update)."""
    if frac >= 1.0:
        return cfg, cfg["N"], "the full workload"
    n = max(int(cfg["N"] * frac), 16 * max(cfg["K"], 1))
    if cfg["name"] == "C4":
        n_old = max(int(cfg["theta_old"].shape[0] * frac), 64)
        s = dict(cfg, theta_new=np.asfortranarray(cfg["theta_new"][:n]), theta_old=np.asfortranarray(cfg["theta_old"][:n_old]), w_old=cfg["w_old"][:n_old])
        return s, n, (f"{n} new x {n_old} old particles of the 1M x 1M update; value = N / (t * (N/{n}) * (N_old/{n_old})) "
                      f"(exact cost law N_new*N_old*P)")
    npp = max(int(cfg["N_pp"] * frac), 16)
    s = dict(cfg, N=n, N_pp=npp, metrics=np.asfortranarray(cfg["metrics"][:n]), params=np.asfortranarray(cfg["params"][:n]))
    return s, n, (f"first {n} of {cfg['N']} particles ranked, {npp} of {cfg['N_pp']} new rows re-weighted against the full previous prior; "
                  f"value = sample particles / sample time (every stage is linear in N up to the log factor of the sorts)")


def oracle_step(cfg, orc):
    """One step of the hot path on the CPU oracle, the reference's call sequence (AbcSmc.cpp:634-664, 1041-1066)."""
    if cfg["name"] == "C4":
        return orc.weight_predictive_prior(np.ones(cfg["theta_new"].shape[0]), cfg["theta_new"], cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
    r = orc.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    sel = cfg["params"][r["order"][:cfg["N_pp"]].astype(np.int64), :]
    orc.calculate_doubled_variance(sel)
    return orc.weight_predictive_prior(np.ones(len(sel)), sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"])


# rough single-thread seconds per full step on a ~3 GHz host core, used only to size the bounded sample
CPU_STEP_SECONDS = {"C2": 1.0, "C3": 210.0, "C4": 4.8e5, "C5": 4000.0, "T1M": 850.0}


def time_oracle(cfg, budget_s, reps_max):
    import oracle as orc
    orc.build()
    full = CPU_STEP_SECONDS.get(cfg["name"], 60.0)
    frac = 1.0 if full <= budget_s else (np.sqrt(budget_s / full) if cfg["name"] == "C4" else budget_s / full)
    sample, n, what = oracle_sample(cfg, frac)
    reps, t0 = 0, time.perf_counter()
    while reps < reps_max and (reps == 0 or time.perf_counter() - t0 < budget_s):
        oracle_step(sample, orc); reps += 1
    dt = (time.perf_counter() - t0) / reps
    if cfg["name"] == "C4":
        n_old = sample["theta_old"].shape[0]
        value = cfg["N"] / (dt * (cfg["N"] / n) * (cfg["theta_old"].shape[0] / n_old))
    else:
        value = n / dt
    return value, dt, reps, what


def reference_sources_sample(cfg, n=10000):
    """The reference's OWN sources (oracle/_ref/libabcref.so: unmodified pls.cpp + AbcUtil.cpp on the Eigen / GSL stand-ins, built in the
    authoring container, travels prebuilt) and the oracle port timed on the same bounded sample, one pass each: the evidence for timing
    the port (the stand-in's products are naive loops, so that build is the slower of the two). Ranking only; None when the library is
    absent or the workload has no ranking."""
    if cfg["name"] == "C4":
        return None
    try:
        import oracle as orc
        import oracle.ref as ref
        if not os.path.exists(ref.LIB_PATH):
            return None
        n = min(n, cfg["N"])
        met, par = np.asfortranarray(cfg["metrics"][:n]), np.asfortranarray(cfg["params"][:n])
        t0 = time.perf_counter(); o_ref = ref.particle_ranking_PLS(met, par, cfg["target"], 0.5); t_ref = time.perf_counter() - t0
        t0 = time.perf_counter(); o_port = orc.particle_ranking_PLS(met, par, cfg["target"], 0.5)["order"]; t_port = time.perf_counter() - t0
        return {"particles": n, "reference_sources_seconds": t_ref, "port_seconds": t_port, "kind": "reference (Eigen / GSL stand-in headers)",
                "orders_identical": bool(np.array_equal(np.asarray(o_ref).astype(np.int64), np.asarray(o_port).astype(np.int64)))}
    except Exception as e:          # the baseline line must not depend on this extra
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = make_workload(args.workload)
    steps = max(1, args.steps)
    full = CPU_STEP_SECONDS.get(cfg["name"], 60.0)
    warm_done = 0
    if full <= 300.0:
        # the FULL workload: as many passes of the `steps` asked for as fit in ~4 minutes, at least one (C3: exactly one, ~200 s)
        reps_max = max(1, min(steps, int(240.0 / full)))
        if args.warmup > 0 and full <= 5.0:
            time_oracle(cfg, 1e9, 1); warm_done = 1
        value, dt, reps, what = time_oracle(cfg, 1e9, reps_max)
    else:
        budget = max(2.0, min(20.0, 150.0 / (steps + min(args.warmup, 1))))     # whole run within a few minutes
        if args.warmup > 0:
            time_oracle(cfg, budget, 1); warm_done = 1
        value, dt, reps, what = time_oracle(cfg, budget * steps, steps)
    ms = cfg["N"] / value * 1e3
    ref_sources = reference_sources_sample(cfg)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": reps, "warmup": warm_done,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if cfg["name"] == "C4" else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(cfg, 1, "reference"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": what,
                             "seconds_per_pass": dt,
                             "note": "oracle/abc_oracle.cpp (restatement, pinned to the reference's own sources compiled with Eigen / GSL stand-in headers: tests/test_ref_pin.py; that build, oracle/_ref, is 1.8x slower than this port because its inner kernels are naive loops, so the port is the fairer baseline); single thread, as the reference runs",
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if ref_sources:
        line["cpu_baseline"]["reference_sources_sample"] = ref_sources
    print(json.dumps(line), flush=True)


# ---- roofline ---------------------------------------------------------------------------------------------------------------
def kernel_work(cfg, world, ctx_stats):
    """Algorithmic bytes / flops per launch of each instrumented kernel (DESIGN.md §5 restates SURVEY.md §8d)."""
    N, P, K, N_pp = cfg["N"], cfg["P"], cfg["K"], cfg["N_pp"]
    N_old = cfg["theta_old"].shape[0]
    w = {}
    if cfg["name"] != "C4":
        n_tr = int(np.floor(N * 0.5 + 0.5)); n_te = N - n_tr
        A = K
        c_used = ctx_stats.get("ncomp_used", K)
        nchk = (A + 3) // 4
        groups = ctx_stats.get("tests", P * (A - 1)) / 4.0
        # The component loop is ONE CTA walking a strict dependency chain (every component needs the deflated XY of the previous one and
        # a dominant eigenvector in between): what it can be held against is the FP64 pipe of the one SM it runs on, not the GPU's.
        # Executed arithmetic per component (fused multiply-adds as 2 flop): S0 = XY^T XY on upper-triangle tiles K M (M + 1); ~10
        # squarings of the M x M iterate 10 M^2 (M + 1); w^ = XY q, q^ = XY^T w^, deflation of XY 6 K M; rank-one update of the packed
        # H and p^ = H w^ 3 K^2. (The wide loop spreads the same arithmetic over three small launches per component: same accounting.)
        loop_flop = A * (K * P * (P + 1) + 10.0 * P * P * (P + 1) + 6.0 * K * P + 3.0 * K * K)
        loop_name = ctx_stats.get("pls_loop", "pls_gram_kernel")
        if loop_name != "pls_defl_kernel":      # three small launches per component (wide shapes) or the L2-streaming one-CTA loop: latency bound, reported against HBM
            w[loop_name] = ("hbm", 8.0 * (K * K + K * P + (4 * K + P) * A), "A strictly sequential components (a dominant eigenvector each): latency bound by construction (profiles/README.md)")
        else:
            w[loop_name] = ("tensor_one_sm", loop_flop,
                           "A strictly sequential components (a dominant eigenvector each) on ONE SM: latency bound by construction; achieved flop/s "
                           "against ONE SM's share of the FP64 DMMA peak (peak / 148) - the other SMs run the kernels that consume its output "
                           "(pipelined fit, DESIGN.md 4); against the whole GPU the fraction is this / 148")
        nTx, nTy = (K + 7) // 8, (P + 7) // 8
        w["gram_kernel"] = ("tensor", 128.0 * n_tr * (nTx * (nTx + 1) // 2 + nTx * nTy),
                            "X^T X (upper triangle) and X^T Y in one pass: 8x8 tile pairs x 128 flop per row, FP64 DMMA")
        # Level 1 / level 2 re-read the same columns many times out of L2 (the 4 score columns of a group serve all P responses, the
        # reference residual of a response all its groups): the HBM figure counts every DISTINCT column once (what DRAM has to deliver;
        # ncu's dram bytes agree, profiles/), the L2-level request stream is reported beside it. Neither bounds the kernels: they are
        # bound by the issue rate of the per-thread counter updates (LDS.U8 / IADD / STS.U8 chains), see DESIGN.md §4.
        lvl2 = ctx_stats.get("level2", 0)
        w["screen1_kernel"] = ("hbm", 8.0 * n_te * (groups + P + A), "distinct columns once: one checkpoint column per group of 4 tests, P reference residuals, A score columns; "
                               f"requests served from L2: {groups * n_te * 48.0:.3e} B per launch (6 loads per row and group); bound by shared-memory counter updates, not by HBM")
        w["screen2_kernel"] = ("hbm", 8.0 * n_te * (lvl2 + P + A), "distinct columns once: a checkpoint column per test reaching level 2, P reference residuals, A score columns; "
                               f"requests served from L2: {lvl2 * n_te * 28.0:.3e} B per launch; bound by shared-memory atomics, not by HBM")
        pb = int(ctx_stats.get("pipe_block", 0))
        if pb:      # pipelined fit + validation: the brackets time the first block of `pb` components (api.cu: rank_fit_holdout_pipelined)
            cb = min(pb, A)
            w["press_chk_kernel"] = ("hbm", 8.0 * n_te * (cb + P + P * (cb // 4)), f"per block of {cb} components: read the block's scores and the starting residuals, write the checkpoints")
            w["xb_kernel<0>"] = ("tensor", 2.0 * N * K * cb, f"scores of ALL N rows for a block of {cb} components, T = Zx R, FP64 DMMA (the projection of AbcUtil.cpp:453-454 is its first c* columns)")
        else:
            w["press_chk_kernel"] = ("hbm", 8.0 * n_te * (A + P + P * max(nchk - 1, 0)), "read T and Y once, write the checkpoints")
            w["xb_kernel<0>"] = ("tensor", 2.0 * n_te * K * A, "hold-out scores T = X_te R, FP64 DMMA")
        inten = c_used / 4.0
        w["xb_kernel<1>"] = (("tensor", 2.0 * N * K * c_used + 3.0 * N * c_used, "projection + distance, FP64 DMMA") if inten > 5.7 else
                             ("hbm", 8.0 * N * K + 8.0 * N, "projection + distance (c*/4 flop/B below the machine balance)"))
        w["zscore_kernel"] = ("hbm", 16.0 * N * K, "read the metrics, write z")
        shard = N_pp
    else:
        shard = (N + world - 1) // world
    w["weights_main_kernel"] = ("tensor", float(shard) * N_old * (3 * P + 2), "(3P+2) flop per pair, exp counted separately (~20 DFMA each)")
    return w


def roofline_objects(cfg, world, kms, ctx_stats):
    hbm, hbm_src = load_hbm_peak()
    f64, f64_src = load_fp64_peak()
    traffic = load_traffic().get(cfg["name"], {})
    work = kernel_work(cfg, world, ctx_stats)
    out = []
    for name, ms in kms.items():
        if ms <= 0 or name not in work:
            continue
        bound, amount, note = work[name]
        if bound == "hbm":
            ach = amount / (ms * 1e-3) / 1e9; peak, unit, src = hbm, "GB/s", hbm_src
        elif bound == "tensor_one_sm":
            bound = "tensor"
            ach = amount / (ms * 1e-3) / 1e12; peak, unit, src = f64 / 148.0, "TFLOP/s", f64_src + " / 148 SMs (one-CTA kernel)"
        else:
            ach = amount / (ms * 1e-3) / 1e12; peak, unit, src = f64, "TFLOP/s", f64_src
        out.append({"kernel": name, "ms_per_launch": ms, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                    "traffic": traffic.get(name), "algorithmic": amount, "note": note, "peak_source": src})
    out.sort(key=lambda d: -d["ms_per_launch"])
    return out


# ---- GPU arm ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C3", choices=["C2", "C3", "C4", "C5", "T1M"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", type=int, default=0, help="0 KERNEL_TYPE1 (reference default), 1 KERNEL_TYPE2, 2 KERNEL_TYPE1 streamed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the C4 sharded weight-update measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2 / C5 / T1M objects of `configs` (N = 1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
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
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=devt)

    def pinned(a):   # numpy (rows, cols) -> pinned column-major host buffer, returned as a Fortran-ordered numpy view
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)).pin_memory()
        return t, t.numpy().T

    def timed(fn, steps, warm, stages=False, kernels=()):
        """W warm-up steps, then `steps` timed ones: barrier + synchronize on both sides, CUDA events on the launching
        stream, max over ranks. Returns (ms per step, per-stage ms, per-kernel ms, launches). `stages` / `kernels` select the
        library's own CUDA-event brackets that are live during these steps (each record is a stream operation between
        launches; at the C2 shape twenty of them cost 14 % of the step, so the timed region carries only the dominant kernel's)."""
        ctx.set_timers(stages=stages, kernels=kernels)
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        total_ms, stage_acc, kern_acc = 0.0, {}, {}
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
            for k, v in ctx.kernel_ms().items():
                kern_acc[k] = kern_acc.get(k, 0.0) + v
        t = torch.tensor([total_ms], dtype=torch.float64, device=devt)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return (float(t.item()) / steps, {k: v / steps for k, v in stage_acc.items()}, {k: v / steps for k, v in kern_acc.items()},
                (ctx.launches - n_launch0) // steps)

    C4_CHECK = {}

    # ---- the sharded weight update on the C4 shape (the whole step when --workload C4) ------------------------------------
    def sharded_weight_update(steps, warm):
        c4 = make_workload("C4")
        t_new, h_new = pinned(c4["theta_new"]); t_old, h_old = pinned(c4["theta_old"])
        d_new = t_new.to(devt)
        d_old = t_old.to(devt) if rank == 0 else torch.empty_like(t_old, device=devt)
        d_w = torch.from_numpy(c4["w_old"]).to(devt) if rank == 0 else torch.empty(c4["w_old"].size, dtype=torch.float64, device=devt)
        d_dv = torch.from_numpy(c4["dv_old"]).to(devt)
        bytes_io = [0, 0]
        # the communicator lives in the C library (sharded.cu: ncclBroadcast of the previous set from rank 0, as north_star states,
        # all-reduce MAX of the conditioning figure, all-reduce SUM of the squared norm, all-gather of the slices)
        sgrp = dev.ShardGroup(ctx) if world > 1 else None

        def step_dev():
            if world > 1:
                return dev.weights_sharded(ctx, None, d_new, d_old, d_w, d_dv, gather=True, shard_group=sgrp, bcast_root=0)
            return dev.weights(ctx, None, d_new, d_old, d_w, d_dv)

        def step_host():
            if world > 1:
                dn = t_new.to(devt, non_blocking=True)
                if rank == 0:
                    d_old.copy_(t_old, non_blocking=True)
                w = dev.weights_sharded(ctx, None, dn, d_old, d_w, d_dv, gather=True, shard_group=sgrp, bcast_root=0)
                bytes_io[0] = (t_new.numel() + (t_old.numel() if rank == 0 else 0)) * 8; bytes_io[1] = w.numel() * 8
                return w.cpu()
            bytes_io[0] = (t_new.numel() + t_old.numel() + c4["w_old"].size + c4["P"]) * 8; bytes_io[1] = c4["N"] * 8
            return api.weight_predictive_prior(None, h_new, h_old, c4["w_old"], c4["dv_old"], ctx=ctx)

        ms_dev, stages, kms, launches = timed(step_dev, steps, warm, stages=True, kernels=("weights_main_kernel",))
        # parity of what was just timed (outside the timed region): 64 rows of the gathered result against (a) the pairwise-difference
        # kernel (the reference's formulation, algo = 1) on the same rows and (b) the CPU oracle's committed output for those rows
        # (tests/golden/fullsize_C4rows.npz); both are L2-normalised over the 64 rows, so the gathered rows are renormalised alike.
        w_full = step_dev()
        if rank == 0:
            gpath = os.path.join(ROOT, "tests", "golden", "fullsize_C4rows.npz")
            rows = ((np.arange(64, dtype=np.int64) * 15625 + 7) % c4["N"])
            ridx = torch.from_numpy(rows).to(devt)
            got = w_full[ridx].cpu().numpy()
            got = got / np.sqrt(np.sum(got * got))
            sub = d_new[:, ridx].contiguous()
            pair = dev.weights(ctx, None, sub, d_old, d_w, d_dv, algo=1).cpu().numpy()
            C4_CHECK["max_rel_err"] = float(np.max(np.abs(got - pair) / np.abs(pair)))
            C4_CHECK["max_rel_err_what"] = "64 rows of the gathered sharded result vs the pairwise-difference kernel (algo=1) on the same rows, both L2-normalised over the 64 rows"
            if os.path.exists(gpath):
                g = np.load(gpath)
                assert np.array_equal(g["rows"], rows)
                C4_CHECK["max_rel_err_vs_oracle"] = float(np.max(np.abs(got - g["w_normalised_over_rows"]) / np.abs(g["w_normalised_over_rows"])))
            C4_CHECK["finite"] = bool(torch.isfinite(w_full).all().item())
        ms_e2e, _, _, _ = timed(step_host, max(1, steps // 2), 1, kernels=("weights_main_kernel",))
        return c4, ms_dev, ms_e2e, stages, kms, launches, bytes_io

    def measure_ranking(name, steps, warm):
        """One ranking workload: device-resident value, end-to-end through the host API, stage / kernel times, roofline entries."""
        cfg = make_workload(name, replica=rank)
        N, P, K, N_pp = cfg["N"], cfg["P"], cfg["K"], cfg["N_pp"]
        t_met, h_met = pinned(cfg["metrics"]); t_par, h_par = pinned(cfg["params"]); t_old, h_old = pinned(cfg["theta_old"])
        d_met, d_par, d_old = t_met.to(devt), t_par.to(devt), t_old.to(devt)
        d_target = torch.from_numpy(cfg["target"]).to(devt)
        d_wold = torch.from_numpy(cfg["w_old"]).to(devt)
        d_dvold = torch.from_numpy(cfg["dv_old"]).to(devt)
        h_target = cfg["target"]
        stats = {}

        def step_device():
            """inputs resident in HBM; outputs stay in HBM"""
            order, _, used, _ = dev.rank_pls(ctx, d_met, d_par, d_target, 0.5, top_n=N_pp, method=args.method)
            stats["ncomp_used"] = used
            g, dv = dev.doubled_variance_gather(ctx, d_par, order)
            return dev.weights(ctx, None, g, d_old, d_wold, d_dvold)

        def step_host():
            """the reference-facing call sequence on host buffers (AbcSmc.cpp:634-664, 1041-1066)"""
            order = api.particle_ranking_PLS(h_met, h_par, h_target, 0.5, top_n=N_pp, method=args.method, ctx=ctx)
            sel = np.asfortranarray(h_par[order.astype(np.int64), :])
            api.calculate_doubled_variance(sel, ctx=ctx)
            return api.weight_predictive_prior(None, sel, h_old, cfg["w_old"], cfg["dv_old"], ctx=ctx)

        h2d_bytes = (N * (K + P) + K + 2 * N_pp * P + h_old.size + cfg["w_old"].size + P) * 8
        d2h_bytes = (N_pp + P + N_pp) * 8
        # (1) instrumented pass, NOT the reported value: every stage and hot kernel bracketed, to find the dominant kernel and to
        #     fill `stages_ms` / `roofline_kernels`; (2) the timed region proper with only the dominant kernel's bracket live.
        _, stages, kms_all, _ = timed(step_device, 3, warm, stages=True, kernels="all")
        dominant = max(kms_all, key=lambda k: kms_all[k])
        ms_dev, _, kms_dom, launches = timed(step_device, steps, 1, kernels=(dominant,))
        kms = dict(kms_all); kms[dominant] = kms_dom[dominant]
        stats["tests"] = ctx.stat(1); stats["level2"] = ctx.stat(2); stats["exact_so_far"] = ctx.stat(3)
        stats["pipe_block"] = ctx.stat(5); stats["sm_partition"] = ctx.stat(6); stats["exact_radix_calls_so_far"] = ctx.stat(7)
        # the ncu name(s) of what timer slot 0 bracketed: the component loop of the PLS fit
        stats["pls_loop"] = {1: "pls_defl_kernel", 3: "wide_s0_kernel + wide_eig_kernel + wide_hw_kernel (x A components)"}.get(ctx.stat(4), "pls_gram_kernel")
        kms = {(stats["pls_loop"] if k == "pls_gram_kernel" else k): v for k, v in kms.items()}
        ms_e2e, stages_e2e, _, launches_e2e = timed(step_host, steps, warm, kernels=(dominant,))
        # the same set through the chained entry point (abcb200_chain_*: one call per set, the previous set restored on the device first;
        # what AbcSmc's set loop would call after the optional change of INTEGRATION.md) - reported beside `e2e`, not instead of it
        chain = api.SmcChain(P, ctx)

        def step_chain():
            chain.restore(h_old, cfg["w_old"], cfg["dv_old"], sets_done=1)
            return chain.process_set(h_met, h_par, h_target, N_pp, filtering=api.FILTER_PLS, training_fraction=0.5, method=args.method, report=True)

        try:
            ms_chain, _, _, _ = timed(step_chain, max(2, steps // 2), 2, kernels=(dominant,))
        finally:
            chain.close()
        res = dict(cfg=cfg, ms_dev=ms_dev, ms_e2e=ms_e2e, ms_chain=ms_chain, stages=stages, stages_e2e=stages_e2e, kms=kms, launches=launches,
                   launches_e2e=launches_e2e, stats=stats, h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes, steps=steps, warm=warm)
        del t_met, t_par, d_met, d_par, h_met, h_par
        torch.cuda.empty_cache()
        return res

    sampler = ClockSampler(local_rank)
    extra = {}
    if args.workload == "C4":
        if rank == 0:
            sampler.start()
        steps = min(args.steps, 3)
        cfg, ms_dev, ms_e2e, stages, kms, launches, bytes_io = sharded_weight_update(steps, min(W, 2))
        units = cfg["N"]
        scaling = "strong"
        stats = {}
        h2d_bytes, d2h_bytes = bytes_io
        launches_e2e = launches
        stages_e2e = {}
        W_used = min(W, 2)
    else:
        steps, W_used = args.steps, W
        if rank == 0:
            sampler.start()
        m = measure_ranking(args.workload, steps, W_used)
        cfg, ms_dev, ms_e2e, stages, stages_e2e, kms = m["cfg"], m["ms_dev"], m["ms_e2e"], m["stages"], m["stages_e2e"], m["kms"]
        launches, launches_e2e, stats, h2d_bytes, d2h_bytes = m["launches"], m["launches_e2e"], m["stats"], m["h2d_bytes"], m["d2h_bytes"]
        units = cfg["N"] * world
        scaling = "weak"
        if world == 1 and not args.no_extra:
            # the other ranking shapes, measured the same way (fewer steps for the big ones); N = 1 only
            for name, st in (("C2", 20), ("C5", 3), ("T1M", 5)):
                if name == args.workload:
                    continue
                x = measure_ranking(name, st, 3)
                xr = roofline_objects(x["cfg"], 1, x["kms"], x["stats"])
                extra[name] = {"workload": workload_config(x["cfg"], 1, "ours")["workload"], "steps": st, "warmup": 3, "ms_per_step": x["ms_dev"],
                               "value": x["cfg"]["N"] / (x["ms_dev"] * 1e-3), "unit": UNIT,
                               "e2e": {"value": x["cfg"]["N"] / (x["ms_e2e"] * 1e-3), "unit": UNIT, "ms_per_step": x["ms_e2e"],
                                       "h2d_bytes_per_step": int(x["h2d_bytes"]), "d2h_bytes_per_step": int(x["d2h_bytes"])},
                               "gpu_launches_per_step": int(x["launches"]), "stages_ms": x["stages"], "selection": x["stats"],
                               "roofline": ({k: xr[0][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "ms_per_launch")} if xr else None),
                               "roofline_kernels": [{k: r[k] for k in ("kernel", "ms_per_launch", "bound", "achieved", "peak", "unit", "frac", "traffic")} for r in xr]}
                x = None

    sharded = None
    if args.workload != "C4" and not args.no_sharded:
        c4, c4_dev, c4_e2e, c4_stages, c4_kms, c4_launches, c4_io = sharded_weight_update(2, 1)
        if rank == 0:
            roofs = roofline_objects(c4, world, c4_kms, {})
            sharded = {"workload": workload_config(c4, world, "ours")["workload"], "n_gpus": world, "steps": 2, "warmup": 1, "ms_per_step": c4_dev,
                       "value": c4["N"] / (c4_dev * 1e-3), "unit": "new particles/s (each against 1M old particles)", "pairs_per_s": c4["N"] * float(c4["theta_old"].shape[0]) / (c4_dev * 1e-3),
                       "scaling": "strong", "e2e_ms_per_step": c4_e2e, "collectives": "NCCL from the C library (abcb200_weights_sharded_dev): broadcast theta_old + w_old + dv_old (248 MB), all-reduce MAX 1 double (kernel choice), all-reduce SUM 1 double, all-gather weights" if world > 1 else "none",
                       "roofline": roofs[0] if roofs else None}
            sharded.update(C4_CHECK)

    clocks = sampler.stop() if rank == 0 else None     # sampled from the first timed step to the end of the last timed region
    if rank == 0:
        roofs = roofline_objects(cfg, world, kms, stats)
        line = {"metric": METRIC, "value": units / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": W_used,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, world, "ours"), "clocks": clocks,
                "e2e": {"value": units / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes)},
                "e2e_chain": ({"value": units / (m["ms_chain"] * 1e-3), "unit": UNIT, "ms_per_step": m["ms_chain"],
                               "what": "host buffers through abcb200_chain_restore + abcb200_chain_process_set (rank, gather, report statistics, doubled variance, weights in one call)"}
                              if args.workload != "C4" else None),
                "gpu_launches": int(launches) * steps, "gpu_launches_per_step": int(launches), "gpu_launches_per_step_e2e": int(launches_e2e),
                "stages_ms": stages, "stages_ms_e2e": stages_e2e, "selection": stats,
                "roofline": ({k: roofs[0][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "ms_per_launch", "note", "peak_source")} if roofs else None),
                "roofline_kernels": roofs,
                "pipeline_note": ("C2 / C3 / T1M: the PLS component loop (one CTA on a high-priority stream; C5: three launches per component on a 24-SM "
                                  "green-context partition, `selection.sm_partition`) runs BESIDE the kernels that consume its "
                                  "output block by block (R columns, scores of all rows, PRESS + checkpoints) on the other SMs, so `stages_ms` pls_fit and "
                                  "holdout_press overlap and their sum exceeds their share of the step; project_distance is a read of the stored scores"),
                "timing_note": ("`roofline` (the dominant kernel) is bracketed by CUDA events inside the timed region; the other entries of "
                                "`roofline_kernels` and `stages_ms` come from a 3-step instrumented pass right before it (all brackets on), "
                                "because every extra event record is a stream operation between launches"),
                "configs": extra,
                "sharded_weight_update": sharded}
        if args.workload == "C4":
            line.update(C4_CHECK)
        if world == 1 and not args.no_cpu_baseline:
            value, dt, reps, what = time_oracle(cfg, 12.0, 3)
            line["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{reps} pass(es) over {what}",
                                    "seconds_per_pass": dt, "host_cores_available": os.cpu_count(),
                                    "note": "oracle/abc_oracle.cpp, g++ -O2, single thread like the reference"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
