#!/usr/bin/env python
"""bench.py — PPO-update samples/s on B200 (metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gemm fp32|tf32x3|f16x3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one filled rollout buffer: returns scan over the whole
buffer (K1) + device permutation (K3) + ONE PPO epoch = ceil(N/B) minibatches of gather (K4) ->
MLP forward (K5) -> fused loss (K6) -> MLP backward (K7) -> [NCCL grad all-reduce] -> Adam (K8).
`value` = transitions processed by all ranks / device time (CUDA events on the library's stream,
max over ranks) with the buffer resident in HBM; `e2e` = the same through the public API with HOST
(pinned) buffers: the H2D append of the whole buffer and the D2H read of the losses are inside the
timed region.  N=1 workload: config C3 (1M transitions, MLP 3x512, B=65536).  N>1: every rank holds
a C3-sized shard (weak scaling; at N=8 this is C4: 8M transitions, global B=65536, B/N rows per
rank) and the minibatch gradients are summed over the ranks inside the Adam kernel through NVLink
peer memory (CUDA IPC; PPO_B200_NO_P2P=1 selects the NCCL all-reduce instead).

--impl reference times the CPU restatement of the reference (oracle/: Julia is not installable in
this image) on a bounded sample of the same workload with all host threads.
"""
from __future__ import annotations

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

METRIC = "ppo_update_samples_per_s"
UNIT = "samples/s"
EPS, W_ENT, ETA, GAMMA = 0.05, 0.01, 1e-4, 1.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_step(cfg, data, W, b, rows, threads):
    """The restated CPU path (oracle port) on `rows` transitions of the workload: serial returns scan
    (C), record-copy gather (C), fp32 MLP through the host BLAS, unfused loss, Adam.  Returns seconds."""
    from oracle import c_oracle as CO
    from oracle import ppo_oracle as O
    t0 = time.perf_counter()
    n = data["reward"].shape[0]
    ret = CO.compute_returns(data["reward"], data["terminal"], GAMMA)
    t_scan = time.perf_counter() - t0
    pol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    pol.W, pol.b = [w.copy() for w in W], [x.copy() for x in b]
    opt = O.Adam(ETA)
    perm1 = CO.feistel_permutation(n, 1) + 1
    t1 = time.perf_counter()
    batch = CO.get_batch(data["feat"], data["mask"], data["action"], data["old"], ret, perm1[:rows])
    feat, mask = batch["state"]
    O.step_batch(pol, opt, feat, mask, batch["selected_action"], batch["selected_action_probability"],
                 batch["returns"], EPS, W_ENT)
    t_batch = time.perf_counter() - t1
    return t_scan * rows / n + t_batch


def host_blas_threads(want_all=False):
    """The numpy BLAS pool does the CPU arm's GEMMs.  torchrun exports OMP_NUM_THREADS=1, so the reference arm asks
    for every core this process may run on; returns (context manager or None, threads in use)."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
    except Exception:        # threadpoolctl missing: report what the environment says
        return None, int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if want_all:
        return threadpool_limits(limits=ncpu), ncpu
    info = [d.get("num_threads", 1) for d in threadpool_info() if d.get("user_api") in ("blas", "openmp")]
    return None, (max(info) if info else 1)


def cpu_baseline(cfg, data, W, b, target_s=12.0):
    _, threads = host_blas_threads()
    rows = min(2048, cfg.B)
    t = cpu_reference_step(cfg, data, W, b, rows, threads)
    rate = rows / t
    rows2 = int(min(cfg.B, max(rows, rate * target_s)))
    if rows2 > rows * 2:
        t = cpu_reference_step(cfg, data, W, b, rows2, threads)
        rows = rows2
    return {"value": rows / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{rows} of {cfg.N} transitions: C scan share + C gather + numpy/BLAS fp32 MLP fwd/bwd + "
                      f"unfused loss + Adam (oracle port; Julia not installable here), {t:.2f} s"}


# ----------------------------------------------------------------------------------------------
def make_data(cfg, P, S, ctx, W, b, rank, cheap_old=False):
    """Synthetic buffer (host, pinned) + old probabilities from the policy's own forward at the
    initial weights, computed by the device path in chunks (outside every timed region)."""
    import torch
    t0 = time.perf_counter()
    import dataclasses
    cfg_r = dataclasses.replace(cfg, cid=cfg.cid + 100 * rank)     # a different shard per rank
    pinned = {}

    def alloc(shape, dtype):
        tt = torch.empty(shape, dtype=torch.from_numpy(np.empty(1, dtype)).dtype, pin_memory=True)
        pinned[len(pinned)] = tt          # keep the pinned storage alive
        return tt.numpy()

    data = S.make_buffer(cfg_r, alloc=alloc)      # generated in place in pinned host memory (no second copy)
    term = alloc(data["terminal"].shape, np.uint8)
    term[:] = data["terminal"]
    data["terminal"] = term
    data["_pins"] = pinned
    sel = np.empty(cfg.N, np.float32)
    if cheap_old:
        sel[:] = 1.0 / np.maximum(1, np.isfinite(data["mask"]).sum(1))
    else:
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
        chunk = 32768
        for s in range(0, cfg.N, chunk):
            e = min(cfg.N, s + chunk)
            pr = P.batch_action_probabilities(pol, P.StateData(data["feat"][s:e], data["mask"][s:e]))
            sel[s:e] = pr[np.arange(e - s), data["action"][s:e] - 1]
        pol.close()
    old = S.make_old_probs(cfg_r, sel)
    po = torch.empty(cfg.N, dtype=torch.float32, pin_memory=True)
    po.numpy()[...] = old
    data["old"] = po.numpy()
    data["_pins"]["old"] = po
    data["_keep"] = None
    log(f"[rank {rank}] synthetic data ready in {time.perf_counter() - t0:.1f} s")
    return data


class _StdoutToStderr:
    """Route fd 1 to stderr while libraries initialise and run (NCCL prints its version banner on stdout), so that
    the only thing rank 0 ever writes to stdout is the JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def run_ours(args):
    with _StdoutToStderr():
        line = _run_ours(args)
    if line is not None:
        print(line, flush=True)


def _run_ours(args):
    import torch
    import torch.distributed as dist
    import ppo_b200 as P
    from ppo_b200 import synthetic as S
    from ppo_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    torch.cuda.set_device(local_rank)
    cfg = S.CONFIGS[args.config]
    B_local = max(1, (args.batch or cfg.B) // world)
    ctx = P.Context(local_rank)
    if world > 1:
        D.init_comm(ctx)
    W, b = S.make_weights(cfg)
    data = make_data(cfg, P, S, ctx, W, b, rank, cheap_old=args.profile)

    gemm_mode = {"fp32": P.GEMM_FP32_SIMT, "tf32x3": P.GEMM_TF32X3_TC, "f16x3": P.GEMM_F16X3_TC}[args.gemm]
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    pol.set_gemm_mode(gemm_mode)
    p2p = world > 1 and os.environ.get("PPO_B200_NO_P2P", "0") != "1" and D.enable_p2p_gradients(pol)
    opt = P.Optimiser(P.Adam(ETA))
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
    stream = torch.cuda.ExternalStream(ctx.stream())

    def fill():
        buf.clear()
        buf.append(data["feat"], data["mask"], data["old"], data["action"], data["reward"], data["terminal"])

    def update(seed):
        P.compute_state_value_(buf, GAMMA)
        return P.ppo_train_(pol, opt, P.construct_dataset(buf), EPS, B_local, 1, W_ENT, seed=seed, out=None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, tag):
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        launches = ctx.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        log(f"[rank {rank}] {tag}: {ms / steps:.2f} ms/step (events), wall {1e3 * wall / steps:.2f} ms/step")
        return ms / steps, clocks, launches

    # ---- device-resident leg -------------------------------------------------------------------
    fill()
    buf.save_rewards()
    last = {}

    def resident_step(i):
        buf.restore_rewards()
        last["loss"] = update(1000 + i)

    ms_step, clocks, launches = timed(resident_step, args.steps, args.warmup, "resident")
    value = cfg.N * world / (ms_step * 1e-3)
    if args.profile:
        line = None
        if rank == 0:
            line = json.dumps({"profile_only": True, "value": value, "ms_per_step": ms_step, "gpu_launches": launches,
                               "launches_total": ctx.launch_count()})
        barrier()
        pol.close(); buf.close(); ctx.close()
        return line

    # ---- end-to-end leg: host buffers in, losses out ----------------------------------------------
    def e2e_step(i):
        fill()
        last["loss"] = update(2000 + i)

    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e, _, _ = timed(e2e_step, e2e_steps, 1, "e2e")
    h2d = cfg.N * (4 * cfg.nf * cfg.nhe + 4 * cfg.A + 8 + 4 + 4 + 1)
    d2h = 16 * ((cfg.N + B_local - 1) // B_local)
    e2e = {"value": cfg.N * world / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e}

    out = None
    if rank == 0:
        pk = peaks()
        # ---- per-kernel rooflines, measured live with CUDA events on the library's stream ----------
        kernels = {}
        M = B_local * cfg.nhe

        def hbm(name, which, n, a=0, b_=0, c=0, iters=10):
            ms, work = ctx.bench_kernel(which, n, a, b_, c, iters, True)
            gbs = work / (ms * 1e-3) / 1e9
            kernels[name] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm"], "unit": "GB/s",
                             "frac": round(gbs / pk["hbm"], 4), "ms": round(ms, 4), "bytes": work}

        hbm("K1_scan_namedN", "scan", cfg.N, 15)
        hbm("K1_scan_64M", "scan", 64 * 1024 * 1024, 15, iters=5)
        hbm("K1K2_scan_64M_with_norm_stats", "scan_norm", 64 * 1024 * 1024, 15, iters=5)
        hbm("K3_shuffle", "shuffle", cfg.N)
        hbm("K4_gather_ldg", "gather0", cfg.N, cfg.nf * cfg.nhe, cfg.A, B_local)
        hbm("K4_gather_bulk", "gather1", cfg.N, cfg.nf * cfg.nhe, cfg.A, B_local)
        hbm("K6_loss_namedB", "loss", B_local, cfg.A)
        hbm("K6_loss_1M", "loss", 1 << 20, cfg.A, iters=5)
        hp = "head16" if args.gemm == "f16x3" else "head"     # the fp16-split engine has its own head kernels
        hbm("K5_head_fwd", hp + "_fwd", M, cfg.H, cfg.apa, iters=5)
        hbm("K7_head_bwd", hp + "_bwd", M, cfg.H, cfg.apa, iters=5)
        hbm("K1_scan_64M_longepisodes", "scan", 64 * 1024 * 1024, 1 << 30, 1, iters=3)
        hbm("K8_adam", "adam", cfg.num_params)
        gname = {"fp32": "gemm", "tf32x3": "tc1", "f16x3": "tc3"}[args.gemm]
        dom = {}
        for kind in ("fwd", "dgrad", "wgrad"):
            ms, flops = ctx.bench_kernel(f"{gname}_{kind}", M, cfg.H, cfg.H, 0, 3, True)
            tf = flops / (ms * 1e-3) / 1e12
            dom[kind] = {"ms": round(ms, 4), "tflops_fp32_equiv": round(tf, 2)}
            kernels[f"K5K7_gemm_{kind}_{cfg.H}x{cfg.H}"] = {
                "bound": "tensor", "achieved": round(tf, 2), "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": round(tf / pk["bf16_sustained"], 4), "ms": round(ms, 4), "flops": flops}
        # the dominant kernel of the step: the hidden-layer forward/dgrad GEMM kernel (tc_gemm_kk_kernel<256>, 50 % of
        # the step in the ncu launch list).  `achieved` counts ALGORITHMIC flops (2 M K N of the fp32 GEMM it
        # replaces); the 3-pass error-compensated scheme issues 3x as many tf32 tensor flops, and tf32 runs at half the
        # bf16 rate, so the ceiling of this scheme is peak / 6.
        fwd = dom["fwd"]
        ach = fwd["tflops_fp32_equiv"]
        passes = 3 if args.gemm in ("tf32x3", "f16x3") else 1
        # ceiling of the 3-pass split: tf32 runs at half the bf16/fp16 rate (peak / 6); fp16 at the full rate (peak / 3)
        ceiling = pk["bf16_sustained"] / (6.0 if args.gemm == "tf32x3" else 3.0)
        kname = {"tf32x3": "tc_gemm_kk_kernel<256>", "f16x3": "f16_gemm_kk_kernel<256, CTA pair>", "fp32": "sgemm_kernel"}[args.gemm]
        roofline = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": round(ach / pk["bf16_sustained"], 4),
                    "traffic": ({"tf32x3": 8.65e9, "f16x3": 4.34e9}.get(args.gemm)
                                if (M == 1 << 20 and cfg.H == 512) else None),
                    "kernel": kname + f" (hidden Dense forward, M={M}, K=N={cfg.H})",
                    "ms_per_launch": fwd["ms"],
                    "peak_source": pk["src"] + " bf16 sustained (MEASURED_PEAKS.json; kernel timed inside a long step)",
                    "tensor_flops_issued_tflops": round(ach * passes, 2),
                    "frac_of_3pass_ceiling": round(ach / ceiling, 4) if passes == 3 else None,
                    "traffic_source": {"tf32x3": "ncu --set full, profiles/r01_ncu_summary.md (dram read 4.38 GB + write 4.27 GB per launch)",
                                       "f16x3": "ncu --set full, profiles/r01_f16_ncu.md (dram read 2.17 GB + write 2.18 GB per launch "
                                                "= the algorithmic bytes of the fp16 hi/lo pairs)"}.get(args.gemm),
                    "detail": dom}
        # the CPU baseline is timed at N = 1 only (torchrun pins OMP_NUM_THREADS=1 and the ranks share the host)
        cpu = cpu_baseline(cfg, data, W, b) if world == 1 else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "f16x3": "f16x3"}[args.gemm],
            "data": "synthetic",
            "config": {"workload": cfg.name + (f" x{world} shards" if world > 1 else ""),
                       "transitions_per_gpu": cfg.N, "minibatch_rows_per_gpu": B_local,
                       "global_minibatch": B_local * world, "mlp": f"{cfg.L}x{cfg.H}", "nf": cfg.nf, "nhe": cfg.nhe,
                       "actions_per_state": cfg.A, "epochs_per_step": 1, "gemm_engine": args.gemm,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "grad_exchange": ("nvlink peer memory, fused into the Adam kernel" if p2p else "nccl all-reduce") if world > 1 else None,
                       "l2": "step inputs (4.6 GB) exceed L2; per-kernel timings flush L2 between launches (write 256 MB, then read 256 MB so that no dirty flush lines are written back inside the timed kernel)"},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "loss_last_step": [float(last["loss"][0][0]), float(last["loss"][1][0])],
        }
    barrier()
    pol.close(); buf.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return json.dumps(out) if out is not None else None


def run_reference(args):
    """The reference arm: the restated CPU path with all host threads on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import ppo_b200  # noqa: F401  (package only for the synthetic generator; no CUDA call is made)
    from ppo_b200 import synthetic as S
    from oracle import ppo_oracle as O
    cfg = S.CONFIGS[args.config]
    n = min(cfg.N, 131072)                      # bounded sample of the buffer
    import dataclasses
    small = dataclasses.replace(cfg, N=n)
    data = S.make_buffer(small)
    W, b = S.make_weights(cfg)
    data["old"] = np.full(n, 1.0 / cfg.A, np.float32)
    limiter, threads = host_blas_threads(want_all=True)      # all host cores, also under torchrun (OMP_NUM_THREADS=1)
    torch.set_num_threads(threads)
    rows = min(cfg.B, 2048)
    t = cpu_reference_step(cfg, data, W, b, rows, threads)
    rows = int(min(cfg.B, n, max(rows, rows / t * 8.0)))   # ~8 s of CPU work per step
    times = []
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue                                         # one warm-up pass is enough on the CPU
        t = cpu_reference_step(cfg, data, W, b, rows, threads)
        if i >= args.warmup:
            times.append(t)
    t = float(np.mean(times))
    v = rows / t
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg.name, "mlp": f"{cfg.L}x{cfg.H}", "sample_rows_per_step": rows},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{rows} transitions per step of config {cfg.name} (restated CPU path; Julia is "
                                      "not installable in this image)"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gemm", default=os.environ.get("PPO_B200_GEMM", "f16x3"), choices=["fp32", "tf32x3", "f16x3"])
    ap.add_argument("--profile", action="store_true",
                    help="only the warm-up + timed resident steps (for ncu launch lists): no data-generation forward, "
                         "no per-kernel hooks, no CPU baseline")
    ap.add_argument("--config", default="c3")
    ap.add_argument("--batch", type=int, default=0,
                    help="override the global minibatch size (profiling the small-minibatch regime of the N-GPU runs on one GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
