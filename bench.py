#!/usr/bin/env python
"""bench.py — PPO-update samples/s on B200 (metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gemm fp32|tf32x3|f16x3]
                    [--config c3] [--batch B] [--scaling weak|strong] [--e2e-feat i8|f32] [--profile] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one filled rollout buffer: returns scan over the whole buffer (K1) + device
permutation (K3) + ONE PPO epoch = ceil(N/B) minibatches of gather (K4) -> MLP forward (K5) -> fused loss (K6) -> MLP
backward (K7) -> [gradient exchange] -> Adam (K8).
`value` = transitions processed by all ranks / device time (CUDA events on the library's stream, max over ranks) with the
buffer resident in HBM; `e2e` = the same through the public API with HOST (pinned) buffers, timed over the same number
of steps: the H2D append of the whole buffer and the D2H read of the losses are inside the timed region.  The host
features are the small integers a quad-game state holds (Matrix{Int64} in the reference), handed over as Int8, the
0 / -Inf action masks as one bit per action, through ppo_buffer_append_packed (--e2e-feat f32 / --e2e-mask f32: as
Float32, 4x / 32x the bytes).
N=1 workload: config C3 (1M transitions, MLP 3x512, B=65536).  N>1: every rank holds a C3-sized shard (weak scaling; at
N=8 this IS config C4: 8M transitions in total, global B=65536, B/N rows per rank); --scaling strong runs C4 as written
at any N (8 388 608 transitions in total, N/G per rank).  The minibatch gradients are summed over the ranks inside the Adam
kernel through NVLink peer memory (CUDA IPC; PPO_B200_NO_P2P=1 selects the NCCL all-reduce instead).

Besides the headline the line carries (each measured live in this run):
  roofline      dominant kernel (hidden Dense forward GEMM): 60 back-to-back launches against the SUSTAINED bf16 peak (the
                regime of the step), roofline.timed_alone against the burst peak; roofline.hbm (scan / gather / loss, each
                timed alone on a settled GPU, against the HBM copy peak); roofline.traffic / roofline.step_traffic from ncu
                replays of that kernel / of one whole step launched by this script (null when ncu is unusable)
  config.token_compaction   the library's default (the MLP skips fully masked tokens) and the same steps with every token
  cpu_baseline  the restated CPU path (oracle port) on a bounded sample, rank 0 at N=1
  configs       the other named configs: C2 (65k transitions, Policy(72,128,2,4)) at B=32 and B=4096, C5 (disk replay)
  dp_parity     N>1: a short sharded epoch after the timed region — replicas bit-identical, losses / weights vs the
                CPU restatement of the sharded scheme on rank 0

--impl reference times the CPU restatement of the reference (oracle/: Julia is not installable in this image) on a
bounded sample of the same workload with all host threads.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ppo_update_samples_per_s"
UNIT = "samples/s"
EPS, W_ENT, ETA, GAMMA = 0.05, 0.01, 1e-4, 1.0
C4_TOTAL = 8388608


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


def workload(args):
    """(cfg of one rank's shard, local minibatch rows, static `config` object) — shared by both arms so that the
    reference arm's line describes exactly the same workload."""
    from ppo_b200 import synthetic as S
    world = max(1, args.gpus)
    cfg = S.CONFIGS[args.config]
    name = cfg.name
    if args.scaling == "strong":
        assert C4_TOTAL % world == 0
        cfg = dataclasses.replace(S.CONFIGS["c4"], N=C4_TOTAL // world)
        name = f"{S.CONFIGS['c4'].name} ({C4_TOTAL} transitions in total, {cfg.N} per GPU)"
    elif world > 1:
        name = f"{cfg.name} x{world} shards" + (" (= config C4: 8M transitions in total)" if world == 8 and args.config == "c3" else "")
    B_local = max(1, (args.batch or cfg.B) // world)
    config = {
        "workload": name, "transitions_per_gpu": cfg.N, "minibatch_rows_per_gpu": B_local,
        "global_minibatch": B_local * world, "mlp": f"{cfg.L}x{cfg.H}", "nf": cfg.nf, "nhe": cfg.nhe,
        "actions_per_state": cfg.A, "epochs_per_step": 1, "gemm_engine": args.gemm,
        "parallelism": f"dp{world}" if world > 1 else "single",
        "grad_exchange": (("nccl all-reduce" if os.environ.get("PPO_B200_NO_P2P", "0") == "1" else
                           "nvlink peer memory, fused into the Adam kernel") if world > 1 else None),
        "host_features": {"i8": "int8", "f32": "float32"}[args.e2e_feat] + " features, " +
                         {"bits": "1-bit action masks (ppo_buffer_append_packed)", "f32": "float32 action masks"}[args.e2e_mask],
        "l2": "step inputs exceed L2; per-kernel timings flush L2 between launches (write 256 MB, then read 256 MB)",
    }
    return cfg, B_local, config


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ----------------------------------------------------------------------------------------------
def cpu_reference_step(cfg, data, W, b, rows):
    """The restated CPU path (oracle port) on `rows` transitions of the workload: serial returns scan (C), record-copy
    gather (C), fp32 MLP through the host BLAS, unfused loss, Adam.  Returns seconds."""
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


def cpu_sample_data(cfg, n_sub):
    """First n_sub transitions of the workload as the oracle's (Float32) arrays, old probabilities drawn by the same
    generator as the GPU arm from the CPU restatement's own forward pass at the initial weights."""
    from ppo_b200 import synthetic as S
    from oracle import ppo_oracle as O
    small = dataclasses.replace(cfg, N=n_sub)
    data = S.make_buffer(small)
    W, b = S.make_weights(cfg)
    pol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    pol.W, pol.b = [w.copy() for w in W], [x.copy() for x in b]
    sel = np.empty(n_sub, np.float32)
    for s in range(0, n_sub, 4096):
        e = min(n_sub, s + 4096)
        pr = O.batch_action_probabilities(pol, data["feat"][s:e], data["mask"][s:e])
        sel[s:e] = pr[np.arange(e - s), data["action"][s:e] - 1]
    data["old"] = S.make_old_probs(small, sel)
    return data, W, b


def cpu_baseline(cfg, target_s=12.0):
    _, threads = host_blas_threads()
    n_sub = min(cfg.N, 65536)
    data, W, b = cpu_sample_data(cfg, n_sub)
    rows = min(2048, cfg.B, n_sub)
    t = cpu_reference_step(cfg, data, W, b, rows)
    rate = rows / t
    rows2 = int(min(cfg.B, n_sub, max(rows, rate * target_s)))
    if rows2 > rows * 2:
        t = cpu_reference_step(cfg, data, W, b, rows2)
        rows = rows2
    return {"value": rows / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one minibatch of {rows} transitions out of the first {n_sub} of config {cfg.name}: C scan share + C "
                      f"gather + numpy/BLAS fp32 MLP fwd/bwd + unfused loss + Adam (oracle port; Julia not installable "
                      f"here), {t:.2f} s"}


# ----------------------------------------------------------------------------------------------
# synthetic data in pinned host memory
# ----------------------------------------------------------------------------------------------
def make_data(cfg, P, S, ctx, W, b, rank, feat_dtype=np.int8, cheap_old=False):
    """Synthetic buffer (host, pinned; features as small integers of `feat_dtype`) + old probabilities from the policy's
    own forward at the initial weights, computed by the device path in chunks (outside every timed region)."""
    import torch
    t0 = time.perf_counter()
    cfg_r = dataclasses.replace(cfg, cid=cfg.cid + 100 * rank)     # a different shard per rank
    pinned = {}

    def alloc(shape, dtype):
        tt = torch.empty(shape, dtype=torch.from_numpy(np.empty(1, dtype)).dtype, pin_memory=True)
        pinned[len(pinned)] = tt          # keep the pinned storage alive
        return tt.numpy()

    data = S.make_buffer(cfg_r, alloc=alloc, feat_dtype=feat_dtype)   # generated in place in pinned host memory
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
            pr = P.batch_action_probabilities(pol, P.StateData(data["feat"][s:e].astype(np.float32), data["mask"][s:e]))
            sel[s:e] = pr[np.arange(e - s), data["action"][s:e] - 1]
        pol.close()
    old = alloc((cfg.N,), np.float32)
    old[...] = S.make_old_probs(cfg_r, sel)
    data["old"] = old
    # the same masks as one bit per action (what ppo_buffer_append_packed takes; a Julia BitMatrix's chunks)
    data["mask_bits"] = P.pack_action_mask(data["mask"], out=alloc((-(-cfg.N * cfg.A // 64),), np.uint64))
    log(f"[rank {rank}] synthetic data for {cfg.name} ready in {time.perf_counter() - t0:.1f} s")
    return data


def host_bytes_per_transition(cfg, feat_dtype, mask_bits=False):
    return np.dtype(feat_dtype).itemsize * cfg.nf * cfg.nhe + (cfg.A / 8.0 if mask_bits else 4 * cfg.A) + 8 + 4 + 4 + 1


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


def bind_to_gpu_numa_node(index):
    """One process per GPU: run this rank (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA
    node its GPU hangs off, like `numactl --cpunodebind --membind` in a production launcher.  With 8 ranks each pulling
    its shard from host memory per e2e step, buffers that all sit on one node halve the aggregate H2D rate.  Returns a
    short description for the log; never fatal."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)      # the CUDA ordinal (honours CUDA_VISIBLE_DEVICES)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(f"{base}/numa_node").read())
        cpus = []
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if node < 0 or not cpus:
            return f"gpu {index}: no NUMA information ({bus})"
        os.sched_setaffinity(0, cpus)
        return f"gpu {index} ({bus}) -> NUMA node {node}, {len(cpus)} cpus"
    except Exception as e:   # noqa: BLE001
        return f"gpu {index}: NUMA binding skipped ({type(e).__name__}: {e})"


class Harness:
    """process-wide handles + the timing loop (barrier + synchronize on both sides, CUDA events on the library's stream,
    max over ranks)"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import ppo_b200 as P
        from ppo_b200 import distributed as D
        from ppo_b200 import synthetic as S
        self.torch, self.dist, self.P, self.D, self.S = torch, dist, P, D, S
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            if os.environ.get("PPO_B200_NO_NUMA_BIND", "0") != "1":
                log(f"[rank {self.rank}] {bind_to_gpu_numa_node(self.local_rank)}")
            dist.init_process_group("cpu:gloo,cuda:nccl", rank=self.rank, world_size=self.world)
        torch.cuda.set_device(self.local_rank)
        self.ctx = P.Context(self.local_rank)
        if self.world > 1:
            D.init_comm(self.ctx)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream())
        self.use_p2p = self.world > 1 and os.environ.get("PPO_B200_NO_P2P", "0") != "1"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, tag, sample_clocks=False):
        torch = self.torch
        for i in range(warmup):
            fn(i)
        self.barrier()
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.ctx.launch_count()
        t0 = time.perf_counter()
        e0.record(self.stream)
        for i in range(steps):
            fn(warmup + i)
        e1.record(self.stream)
        self.barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = self.ctx.launch_count() - l0
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        log(f"[rank {self.rank}] {tag}: {ms / steps:.3f} ms/step (events), wall {1e3 * wall / steps:.3f} ms/step")
        return ms / steps, clocks, launches


def run_update_legs(h, cfg, B_local, gemm_mode, data, W, b, steps, warmup, feat_key="feat", sample_clocks=True, e2e=True,
                    compact=True, mask_key="mask"):
    """resident + end-to-end legs of one workload; returns a dict of raw measurements.  compact: token compaction (the
    library's default: the MLP skips tokens all of whose actions are masked; exact zeros either way)"""
    P = h.P
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, h.ctx, weights=W, biases=b, gemm_mode=gemm_mode)
    pol.set_token_compaction(compact)
    if h.use_p2p:
        h.D.enable_p2p_gradients(pol)
    opt = P.Optimiser(P.Adam(ETA))
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, h.ctx)
    last = {}

    def fill():
        buf.clear()
        buf.append(data[feat_key], data[mask_key], data["old"], data["action"], data["reward"], data["terminal"])

    def update(seed):
        P.compute_state_value_(buf, GAMMA)
        return P.ppo_train_(pol, opt, P.construct_dataset(buf), EPS, B_local, 1, W_ENT, seed=seed, out=None)

    fill()
    buf.save_rewards()

    def resident_step(i):
        buf.restore_rewards()
        last["loss"] = update(1000 + i)

    ms_step, clocks, launches = h.timed(resident_step, steps, warmup, f"{cfg.name} resident", sample_clocks)
    out = {"ms_step": ms_step, "clocks": clocks, "launches": launches, "engine": pol.gemm_mode,
           "nbatches": (cfg.N + B_local - 1) // B_local, "active_tokens_last_minibatch": pol.active_tokens()}
    if h.use_p2p:
        ns, waits = pol.p2p_wait(reset=True)
        out["p2p_wait_us_per_minibatch"] = round(ns / max(1, waits) / 1e3, 2)
    if e2e:
        def e2e_step(i):
            fill()
            last["loss"] = update(2000 + i)

        ms_e2e, _, _ = h.timed(e2e_step, steps, max(1, min(warmup, 2)), f"{cfg.name} e2e")
        out["ms_e2e"] = ms_e2e
    out["loss"] = [float(last["loss"][0][0]), float(last["loss"][1][0])]
    pol.close(); buf.close()
    return out


# ----------------------------------------------------------------------------------------------
# extra legs
# ----------------------------------------------------------------------------------------------
def kernel_rooflines(h, cfg, B_local, gemm):
    """per-kernel rooflines, measured live with CUDA events on the library's stream (rank 0)"""
    pk = peaks()
    ctx = h.ctx
    kernels = {}
    M = B_local * cfg.nhe
    # The timed legs before this left the GPU at its power-cap clocks, and the governor keeps them there for a few hundred
    # milliseconds: a kernel "timed alone" right after them is really timed at the throttled clock (scripts/
    # scan_after_load.py: the 64 M-transition scan takes 215 us in the first ~200 ms after a second of back-to-back GEMMs,
    # 118 us cold and from then on).  Kernels timed alone are therefore timed on a settled GPU and compared with the
    # BURST peaks; the GEMMs are also timed back to back (the regime of the step) and compared with the SUSTAINED peak.
    time.sleep(2.0)

    def hbm(name, which, n, a=0, b_=0, c=0, iters=10):
        ms, work = ctx.bench_kernel(which, n, a, b_, c, iters, True)
        gbs = work / (ms * 1e-3) / 1e9
        kernels[name] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm"], "unit": "GB/s",
                         "frac": round(gbs / pk["hbm"], 4), "ms": round(ms, 4), "bytes": work}
        return kernels[name]

    hbm("K1_scan_namedN", "scan", cfg.N, 15)
    # (20 launches each: a 0.13 ms kernel timed 5 times lets one hiccup of the box halve the reported fraction)
    k1 = hbm("K1_scan_64M", "scan", 64 * 1024 * 1024, 15, iters=20)
    k1g = hbm("K1_scan_64M_gamma0.99", "scan", 64 * 1024 * 1024, 15, 1, iters=20)
    k2 = hbm("K1K2_scan_64M_with_norm_stats", "scan_norm", 64 * 1024 * 1024, 15, iters=20)
    # stress variants: episodes as long as a whole scan tile (4 096 transitions: the look-ahead window
    # misses, every tile looks back one or two tiles), and ONE unterminated episode over all 64 M transitions (every
    # tile's carry depends on every tile to its right: the decoupled look-back's worst case)
    k1s = hbm("K1_scan_64M_episodes4096_gamma0.99", "scan", 64 * 1024 * 1024, 4096, 1, iters=3)
    k1l = hbm("K1_scan_64M_longepisodes_gamma0.99", "scan", 64 * 1024 * 1024, 1 << 30, 1, iters=3)
    hbm("K3_shuffle", "shuffle", cfg.N)
    k4 = hbm("K4_gather_ldg", "gather0", cfg.N, cfg.nf * cfg.nhe, cfg.A, B_local, iters=20)
    hbm("K4_gather_bulk", "gather1", cfg.N, cfg.nf * cfg.nhe, cfg.A, B_local)
    hbm("K6_loss_namedB", "loss", B_local, cfg.A)
    k6 = hbm("K6_loss_1M", "loss", 1 << 20, cfg.A, iters=20)
    hp = "head16" if gemm == "f16x3" else "head"     # the fp16-split engine has its own head kernels
    hbm("K5_head_fwd", hp + "_fwd", M, cfg.H, cfg.apa, iters=5)
    hbm("K7_head_bwd", hp + "_bwd", M, cfg.H, cfg.apa, iters=5)
    hbm("K8_adam", "adam", cfg.num_params)
    gname = {"fp32": "gemm", "tf32x3": "tc1", "f16x3": "tc3"}[gemm]
    dom = {}
    time.sleep(1.0)
    for kind in ("fwd", "dgrad", "wgrad"):          # timed alone: 3 launches, L2 flushed, settled GPU -> burst peak
        ms, flops = ctx.bench_kernel(f"{gname}_{kind}", M, cfg.H, cfg.H, 0, 3, True)
        dom[kind] = {"isolated_ms": round(ms, 4), "isolated_tflops_fp32_equiv": round(flops / (ms * 1e-3) / 1e12, 2),
                     "isolated_frac_of_burst_peak": round(flops / (ms * 1e-3) / 1e12 / pk["bf16"], 4)}
        time.sleep(0.5)
    for kind in ("fwd", "dgrad", "wgrad"):          # back to back (operands >> L2): the regime of the step -> sustained peak
        ms, flops = ctx.bench_kernel(f"{gname}_{kind}", M, cfg.H, cfg.H, 0, 60, False)
        tf = flops / (ms * 1e-3) / 1e12
        dom[kind].update({"ms": round(ms, 4), "tflops_fp32_equiv": round(tf, 2)})
        kernels[f"K5K7_gemm_{kind}_{cfg.H}x{cfg.H}"] = {
            "bound": "tensor", "achieved": round(tf, 2), "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
            "frac": round(tf / pk["bf16_sustained"], 4), "ms": round(ms, 4), "flops": flops,
            "timing": "60 back-to-back launches (sustained clocks)", "isolated_ms": dom[kind]["isolated_ms"]}
    # The dominant kernel of the step: the hidden-layer forward/dgrad GEMM kernel (about half of the step in the ncu
    # launch list).  `achieved` counts ALGORITHMIC flops (2 M K N of the fp32 GEMM it replaces); the error-compensated
    # split issues 3x as many tensor flops (fp16 at the full bf16 rate, tf32 at half of it), so the scheme's own ceiling is
    # peak / 3 (fp16) or peak / 6 (tf32).
    fwd = dom["fwd"]
    ach = fwd["tflops_fp32_equiv"]
    passes = 3 if gemm in ("tf32x3", "f16x3") else 1
    ceiling = pk["bf16_sustained"] / (6.0 if gemm == "tf32x3" else 3.0)
    kname = {"tf32x3": "tc_gemm_kk_kernel<256>", "f16x3": "f16_gemm_kk_kernel<256, CTA pair>", "fp32": "sgemm_kernel"}[gemm]
    roofline = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["bf16_sustained"], 4), "traffic": None,
                "kernel": kname + f" (hidden Dense forward, M={M}, K=N={cfg.H})", "ms_per_launch": fwd["ms"],
                "peak_source": pk["src"] + " bf16 sustained (MEASURED_PEAKS.json); the kernel is timed over 60 back-to-back "
                               "launches, the regime it runs in inside the step",
                "timed_alone": {"ms_per_launch": fwd["isolated_ms"], "achieved": fwd["isolated_tflops_fp32_equiv"],
                                "peak": pk["bf16"], "frac": fwd["isolated_frac_of_burst_peak"],
                                "peak_source": pk["src"] + " bf16 burst; 3 launches, L2 flushed, settled GPU"},
                "tensor_flops_issued_tflops": round(ach * passes, 2),
                "frac_of_3pass_ceiling": round(ach / ceiling, 4) if passes == 3 else None,
                # the HBM-bound kernels of the path against the measured copy peak (>> L2 sizes; named sizes in `kernels`);
                # each timed alone on a settled GPU
                "hbm": {"peak_gbs": pk["hbm"], "K1_scan_64M": k1["frac"], "K1_scan_64M_gamma0.99": k1g["frac"],
                        "K1K2_scan_with_norm_stats": k2["frac"], "K1_scan_64M_episodes_of_4096": k1s["frac"],
                        "K1_scan_64M_one_episode": k1l["frac"],
                        "K4_gather": k4["frac"], "K6_loss_1M": k6["frac"]},
                "detail": dom}
    return roofline, kernels


def ncu_traffic_probe(M, H, which="tc3_fwd"):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel at the benchmark's shape, from an
    ncu replay launched here (a separate tiny process: the timed numbers of this run are never taken under a profiler)"""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); import ppo_b200 as P; c = P.Context(0); "
            f"c.bench_kernel({which!r}, {M}, {H}, {H}, 0, 1, False); c.close()")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:f16_gemm_kk|tc_gemm_kk|sgemm", "-c", "1", "--csv", sys.executable, "-c", code]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    except Exception as e:   # noqa: BLE001
        return None, f"ncu failed: {e}"
    total, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    found = 0
    for ln in r.stdout.splitlines():
        if "dram__bytes_" not in ln:
            continue
        f = [x.strip('"') for x in ln.strip().split('","')]
        try:
            total += float(f[-1].replace(",", "")) * unit_scale.get(f[-2], 1.0)
            found += 1
        except (ValueError, IndexError):
            continue
    if found < 2:
        return None, "ncu produced no dram__bytes rows: " + (r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "empty output")
    return total, "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum replay of one launch, run by bench.py"


def ncu_step_traffic_probe(args, launches_per_step):
    """Measured DRAM traffic of ONE WHOLE STEP (returns scan + permutation + every minibatch of the epoch): this script is
    replayed with `--profile --steps 1 --warmup 1` under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` in a
    separate process (one pass per kernel; no timing is taken from it) and the per-kernel bytes are summed over a window of
    exactly one step's launches (the step is periodic, so any window of that length behind the one-off buffer append
    holds every kernel of a step once)."""
    import csv
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    with tempfile.TemporaryDirectory() as td:
        log_csv = os.path.join(td, "step.csv")
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv", "--log-file",
               log_csv, "--launch-skip", "12", "--launch-count", str(int(launches_per_step)), sys.executable, os.path.abspath(__file__), "--profile", "--steps", "1", "--warmup", "1", "--config",
               args.config, "--gemm", args.gemm, "--e2e-feat", args.e2e_feat, "--e2e-mask", args.e2e_mask]
        if args.batch:
            cmd += ["--batch", str(args.batch)]
        env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
        try:
            subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
            rows = list(csv.reader(open(log_csv)))
        except Exception as e:   # noqa: BLE001
            return {"error": f"ncu step probe failed: {e}"[:200]}
    hdr = next((i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "ID"), None)
    if hdr is None:
        return {"error": "ncu step probe: no launch list"}
    H = rows[hdr]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    append_only = ("narrow_to_f32", "mask_from_bits", "actions_in", "normalize_bool", "absmax_f32", "i64_to_f32")
    rd = wr = 0.0
    launches = 0
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        d = dict(zip(H, r))
        if any(k in d["Kernel Name"] for k in append_only):
            continue
        try:
            v = float(d["Metric Value"].replace(",", "")) * scale.get(d["Metric Unit"], 1.0)
        except ValueError:
            continue
        if d["Metric Name"].startswith("dram__bytes_read"):
            rd += v
            launches += 1
        elif d["Metric Name"].startswith("dram__bytes_write"):
            wr += v
    if launches == 0:
        return {"error": "ncu step probe: no dram__bytes rows"}
    return {"dram_read_gb_per_step": round(rd / 1e9, 2), "dram_write_gb_per_step": round(wr / 1e9, 2),
            "kernel_launches_profiled": launches,
            "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over `bench.py --profile --steps 1 --warmup 1`, "
                      "run by bench.py in a separate process (append kernels excluded)"}


def run_c2(h, args, gemm_mode):
    """config C2: 65 536 transitions, Policy(72,128,2,4) on 64 half-edges (test/test_square_mesh.jl:29), at the
    reference-like tiny minibatch and at B = 4096; one GPU"""
    P, S = h.P, h.S
    cfg = S.CONFIGS["c2"]
    W, b = S.make_weights(cfg)
    data = make_data(cfg, P, S, h.ctx, W, b, 0, np.int8)
    out = {}
    steps = max(1, min(args.steps, 5))
    for B in (32, 4096):
        r = run_update_legs(h, cfg, B, gemm_mode, data, W, b, steps, 2, sample_clocks=False,
                            mask_key="mask_bits" if args.e2e_mask == "bits" else "mask")
        nb = r["nbatches"]
        out[f"c2_b{B}"] = {"value": round(cfg.N / (r["ms_step"] * 1e-3), 1), "e2e": round(cfg.N / (r["ms_e2e"] * 1e-3), 1),
                           "ms_per_step": round(r["ms_step"], 3), "us_per_minibatch": round(1e3 * r["ms_step"] / nb, 2),
                           "minibatches": nb, "launches_per_minibatch": round(r["launches"] / (steps * nb), 1),
                           "steps": steps, "engine": r["engine"]}
    # the largest byte mover at C2's shapes (18 KB of features per sample): the gather, against the HBM peak
    ms, work = h.ctx.bench_kernel("gather0", cfg.N, cfg.nf * cfg.nhe, cfg.A, 4096, 10, True)
    out["c2_b4096"]["K4_gather_frac_of_hbm"] = round(work / (ms * 1e-3) / 1e9 / peaks()["hbm"], 4)
    return out


def run_c5(h, args, gemm_mode):
    """config C5: rollouts written to disk in the reference's format (trajectory.csv + one states/sample_i.bson per
    transition: src/rollouts_to_disk.jl:23-132), bulk-loaded by the C++ replay loader, replicated in the device buffer
    and run through the device update; every rank writes and replays its own shard.  Episodes have random_quad's shapes
    (nf = 216, 4 actions per half-edge, <= 30 steps, Policy(216,128,2,4): test/random_quad.jl:43-49,61); the mesh
    environments themselves need un-vendored packages."""
    P, D = h.P, h.D
    nf, nhe, apa, H, L = 216, 16, 4, 128, 2
    on_disk, replayed = 2048, 131072
    rng = np.random.default_rng(20260118 + 5 + 1000 * h.rank)
    root = tempfile.mkdtemp(prefix=f"c5_rank{h.rank}_")
    try:
        return _run_c5_body(h, gemm_mode, root, rng, nf, nhe, apa, H, L, on_disk, replayed)
    finally:
        shutil.rmtree(root, ignore_errors=True)


def _all_ranks_ok(h, ok):
    """the disk phase is rank-local; the update phase is collective: go on only when every rank got there"""
    if h.world == 1:
        return ok
    t = h.torch.tensor([1 if ok else 0], dtype=h.torch.int32)
    h.dist.all_reduce(t, op=h.dist.ReduceOp.MIN)
    return bool(t.item())


def _run_c5_body(h, gemm_mode, root, rng, nf, nhe, apa, H, L, on_disk, replayed):
    P, D = h.P, h.D
    err = None
    try:
        t0 = time.perf_counter()
        disk = P.DiskRollouts(root)
        host = {"feat": [], "mask": [], "act": [], "prob": []}
        while len(disk) < on_disk:
            steps = int(min(rng.integers(1, 31), on_disk - len(disk)))
            for s in range(steps):
                vs = rng.integers(-3, 9, (nhe, nf)).astype(np.int64)               # Matrix{Int64}[nf, nhe]
                am = np.repeat(np.where(rng.random(nhe // 4) < 0.25, -np.inf, 0.0), 4 * apa).astype(np.float32)
                am[:4 * apa] = 0.0
                a = int(rng.choice(np.flatnonzero(np.isfinite(am)))) + 1
                p = float(np.float32(rng.uniform(0.05, 1.0)))
                P.update_(disk, P.StateData(vs, am), p, a, float(rng.integers(-4, 5)), s == steps - 1)
                host["feat"].append(vs.astype(np.int8)); host["mask"].append(am); host["act"].append(a); host["prob"].append(p)
        P.write_returns_to_disk(disk, 1.0, h.ctx)
        t_write = time.perf_counter() - t0
        n_disk = len(disk)
        t0 = time.perf_counter()
        ds = P.DiskDataset(root)
        buf, has_returns = ds.to_device(nf, nhe, apa, h.ctx, capacity=replayed)
        h.ctx.sync()
        t_load = time.perf_counter() - t0
        got = buf.read(0, n_disk)
        feat, mask = np.stack(host["feat"]), np.stack(host["mask"])
        act, prob = np.array(host["act"], np.int64), np.array(host["prob"], np.float32)
        bit_exact = bool(has_returns and np.array_equal(got["feat"], feat.astype(np.float32)) and np.array_equal(got["mask"], mask)
                         and np.array_equal(got["selected_actions"], act)
                         and np.allclose(got["selected_action_probabilities"], prob, rtol=1e-6))
        ret = buf.rewards.copy()
        while len(buf) + n_disk <= replayed:                                # replicate (returns are already final)
            buf.append(feat, mask, prob, act, ret, np.zeros(n_disk, bool))
        n = len(buf)
    except Exception as e:   # noqa: BLE001
        err = e
    if not _all_ranks_ok(h, err is None):
        raise RuntimeError(f"disk phase failed on a rank: {err}")
    pol = P.Policy(nf, H, L, apa, h.ctx, gemm_mode=gemm_mode)
    if h.use_p2p:
        D.enable_p2p_gradients(pol)
    opt = P.Optimiser(P.Adam(ETA))
    B = max(1, 8192 // h.world)
    dsd = P.construct_dataset(buf)
    epochs = 3

    def step(i):
        P.ppo_train_(pol, opt, dsd, EPS, B, 1, W_ENT, seed=10 + i, out=None)

    ms, _, _ = h.timed(step, epochs, 2, "c5 update")
    out = {"on_disk_per_gpu": n_disk, "replayed_per_gpu": n, "replayed_total": n * h.world, "write_s": round(t_write, 2),
           "load_s": round(t_load, 4), "load_transitions_per_s": round(n_disk / t_load, 1), "loaded_bit_exact": bit_exact,
           "value": round(n * h.world / (ms * 1e-3), 1), "ms_per_epoch": round(ms, 3), "global_minibatch": B * h.world,
           "policy": f"Policy({nf},{H},{L},{apa})", "engine": pol.gemm_mode}
    pol.close(); buf.close()
    return out


def run_dp_parity(h, gemm_mode):
    """N>1: one sharded epoch at C3's widths after the timed region, on the engine / gradient exchange the bench timed
    (CUDA-graph replay included): replicas must be bit-identical, losses and post-Adam weights are compared with the CPU
    restatement of the sharded scheme (global minibatch k = union of every rank's local minibatch k) on rank 0."""
    P, S, D, dist = h.P, h.S, h.D, h.dist
    from oracle import ppo_oracle as O
    cfg = S.Config("c3-widths-dp-parity", 97, 8192, 64, 16, 4, 512, 3, 1024)
    world, rank = h.world, h.rank
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    old = S.rng_for(cfg, 7).uniform(0.05, 1.0, cfg.N).astype(np.float32)
    bounds = D.shard_bounds_at_episode_ends(data["terminal"], world)
    n_use = D.equalize_counts(bounds)
    B = cfg.B // world
    n_use -= n_use % B                       # full minibatches only: every rank replays the same CUDA graph
    a0 = bounds[rank][0]
    sl = slice(a0, a0 + n_use)
    returns = O.compute_returns(data["reward"], data["terminal"], 1.0)      # shards are cut at episode ends
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, n_use, h.ctx)
    buf.append(data["feat"][sl], data["mask"][sl], old[sl], data["action"][sl], returns[sl], data["terminal"][sl])
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, h.ctx, weights=W, biases=b, gemm_mode=gemm_mode)
    if h.use_p2p:
        D.enable_p2p_gradients(pol)
    losses = P.step_epoch_(pol, P.Adam(ETA), P.construct_dataset(buf), EPS, B, W_ENT, seed=D.local_seed(99, rank))
    Wd, bd = pol.weights()
    flat = np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(Wd, bd)])
    gathered = [None] * world
    dist.all_gather_object(gathered, (losses, flat.tobytes()))
    pol.close(); buf.close()
    if rank != 0:
        return None
    flats = [np.frombuffer(g[1], np.float32) for g in gathered]
    identical = all(np.array_equal(flats[0], f) for f in flats[1:]) and all(g[0] == gathered[0][0] for g in gathered[1:])
    opol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    opol.W, opol.b = [w.copy() for w in W], [x.copy() for x in b]
    oopt = O.Adam(ETA)
    perms = [O.feistel_permutation(n_use, D.local_seed(99, r)) + bounds[r][0] for r in range(world)]
    ph, eh = [], []
    for start in range(0, n_use, B):
        idx = np.concatenate([pm[start:start + B] for pm in perms])
        pl, ew = O.step_batch(opol, oopt, data["feat"][idx], data["mask"][idx], data["action"][idx], old[idx],
                              returns[idx], EPS, W_ENT)
        ph.append(pl); eh.append(ew)
    want = (float(np.mean(ph)), float(np.mean(eh)))
    loss_rel = max(abs(losses[0] - want[0]) / abs(want[0]), abs(losses[1] - want[1]) / abs(want[1]))
    dw = np.abs(flats[0] - opol.flat())
    w_err = float(np.max(dw))
    w_p999 = float(np.quantile(dw, 0.999))
    frac_above = float(np.mean(dw > 2e-5))
    disp = float(np.max(np.abs(opol.flat() - np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(W, b)]))))
    steps = n_use // B
    # Adam normalises every step to ~eta whatever the gradient's size, so an entry whose gradient sits at Float32
    # summation-noise level may step either way in two correct evaluations: 99.9 % of the entries within 2e-5, the rest
    # within what the optimiser can move a weight at all (2 eta per minibatch)
    ok = bool(identical and loss_rel <= 1e-5 and frac_above <= 1e-3 and w_err <= 2 * ETA * steps)
    return {"ok": ok, "replicas_bit_identical": bool(identical), "loss_rel_err_vs_sharded_oracle": float(f"{loss_rel:.3g}"),
            "weights_max_abs_err": float(f"{w_err:.3g}"), "weights_p99.9_abs_err": float(f"{w_p999:.3g}"),
            "weights_fraction_above_2e-5": float(f"{frac_above:.3g}"), "weights_max_displacement": float(f"{disp:.3g}"),
            "minibatches": steps, "rows_per_rank": n_use, "mlp": "3x512",
            "tolerances": "loss 1e-5 rel; weights: 99.9 % of entries within 2e-5 abs, all within 2 eta per minibatch"}


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    with _StdoutToStderr():
        line = _run_ours(args)
    if line is not None:
        print(line, flush=True)


def _run_ours(args):
    h = Harness(args)
    P, S = h.P, h.S
    world, rank = h.world, h.rank
    assert world == max(1, args.gpus) or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    cfg, B_local, config = workload(_with_world(args, world))
    feat_dtype = {"i8": np.int8, "f32": np.float32}[args.e2e_feat]
    gemm_mode = {"fp32": P.GEMM_FP32_SIMT, "tf32x3": P.GEMM_TF32X3_TC, "f16x3": P.GEMM_F16X3_TC}[args.gemm]
    W, b = S.make_weights(cfg)
    data = make_data(cfg, P, S, h.ctx, W, b, rank, feat_dtype, cheap_old=args.profile)

    mask_key = "mask_bits" if args.e2e_mask == "bits" else "mask"
    main = run_update_legs(h, cfg, B_local, gemm_mode, data, W, b, args.steps, args.warmup, e2e=not args.profile,
                           mask_key=mask_key)
    value = cfg.N * world / (main["ms_step"] * 1e-3)
    if args.profile:
        line = None
        if rank == 0:
            line = json.dumps({"profile_only": True, "value": value, "ms_per_step": main["ms_step"],
                               "gpu_launches": main["launches"], "launches_total": h.ctx.launch_count(),
                               "p2p_wait_us_per_minibatch": main.get("p2p_wait_us_per_minibatch")})
        h.barrier()
        h.ctx.close()
        return line
    e2e = {"value": cfg.N * world / (main["ms_e2e"] * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(cfg.N * host_bytes_per_transition(cfg, feat_dtype, args.e2e_mask == "bits")),
           "d2h_bytes_per_step": 16 * main["nbatches"], "ms_per_step": main["ms_e2e"], "steps": args.steps}
    # token compaction is the library default; the same steps with every token pushed through the MLP (what the
    # reference does with fully masked tokens) are timed beside it
    if "p2p_wait_us_per_minibatch" in main:
        config["p2p_wait_us_per_minibatch_rank0"] = main["p2p_wait_us_per_minibatch"]
    last_rows = (cfg.N - (main["nbatches"] - 1) * B_local) * cfg.nhe
    config["token_compaction"] = {
        "enabled": main["active_tokens_last_minibatch"] >= 0,
        "active_token_fraction_last_minibatch": round(main["active_tokens_last_minibatch"] / last_rows, 4),
        "mask": "synthetic.make_masks: groups of 4 tokens, inactive with p = 0.25, first group active (SURVEY 8(d))"}
    if args.gemm == "f16x3" and not args.no_extras:
        dense = run_update_legs(h, cfg, B_local, gemm_mode, data, W, b, min(args.steps, 3), 3, sample_clocks=False,
                                e2e=False, compact=False)
        config["token_compaction"]["value_with_every_token"] = cfg.N * world / (dense["ms_step"] * 1e-3)
        config["token_compaction"]["ms_per_step_with_every_token"] = dense["ms_step"]
    del data

    extras, errors = {}, {}

    def guarded(name, fn):
        try:
            return fn()
        except Exception as e:   # noqa: BLE001  (an extra leg must never cost the headline line)
            errors[name] = f"{type(e).__name__}: {e}"[:300]
            log(f"[rank {rank}] extra leg {name} failed: {errors[name]}")
            return None

    roofline, kernels, cpu = None, None, None
    if not args.no_extras:
        if rank == 0:
            res = guarded("kernels", lambda: kernel_rooflines(h, cfg, B_local, args.gemm))
            if res:
                roofline, kernels = res
                if world == 1:
                    tr = guarded("traffic", lambda: ncu_traffic_probe(B_local * cfg.nhe, cfg.H, {"fp32": "gemm_fwd", "tf32x3": "tc1_fwd", "f16x3": "tc3_fwd"}[args.gemm]))
                    if tr:
                        roofline["traffic"], roofline["traffic_source"] = tr
                    roofline["step_traffic"] = guarded("step_traffic", lambda: ncu_step_traffic_probe(args, main["launches"] / args.steps))
            if world == 1:
                # the CPU baseline is timed at N = 1 only (torchrun pins OMP_NUM_THREADS=1 and the ranks share the host)
                cpu = guarded("cpu_baseline", lambda: cpu_baseline(cfg))
                c2 = guarded("c2", lambda: run_c2(h, args, gemm_mode))
                if c2:
                    extras.update(c2)
        c5 = guarded("c5", lambda: run_c5(h, args, gemm_mode))
        if c5:
            extras["c5"] = c5
    dp_parity = guarded("dp_parity", lambda: run_dp_parity(h, gemm_mode)) if world > 1 else None

    out = None
    if rank == 0:
        if roofline is not None:
            # step level: algorithmic MLP flops per step against the measured sustained tensor peak
            pk = peaks()
            step_tf = cfg.flops_per_sample() * cfg.N / (main["ms_step"] * 1e-3) / 1e12
            roofline["step"] = {"algorithmic_tflops": round(step_tf, 1), "frac_of_peak": round(step_tf / pk["bf16_sustained"], 4),
                                "frac_of_3pass_ceiling": round(3 * step_tf / pk["bf16_sustained"], 4)}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "f16x3": "f16x3"}[args.gemm],
            "data": "synthetic", "kernels": kernels, "config": config, "cpu_baseline": cpu,
            "gpu_launches": main["launches"], "clocks": main["clocks"], "loss_last_step": main["loss"],
            "roofline": roofline, "e2e": e2e, "errors": errors or None, "configs": extras or None, "dp_parity": dp_parity,
        }
    h.barrier()
    h.ctx.close()
    if world > 1:
        h.dist.destroy_process_group()
    return json.dumps(out) if out is not None else None


def _with_world(args, world):
    a = argparse.Namespace(**vars(args))
    a.gpus = world
    return a


def run_reference(args):
    """The reference arm: the restated CPU path with all host threads on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import ppo_b200  # noqa: F401  (package only for the synthetic generator; no CUDA call is made)
    cfg, B_local, config = workload(args)
    n = min(cfg.N, 65536)                      # bounded sample of the buffer
    data, W, b = cpu_sample_data(cfg, n)
    limiter, threads = host_blas_threads(want_all=True)      # all host cores, also under torchrun (OMP_NUM_THREADS=1)
    torch.set_num_threads(threads)
    rows = min(cfg.B, 2048, n)
    t = cpu_reference_step(cfg, data, W, b, rows)
    rows = int(min(cfg.B, n, max(rows, rows / t * 8.0)))   # ~8 s of CPU work per step
    times = []
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue                                         # one warm-up pass is enough on the CPU
        t = cpu_reference_step(cfg, data, W, b, rows)
        if i >= args.warmup:
            times.append(t)
    t = float(np.mean(times))
    v = rows / t
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"each step = one minibatch of {rows} transitions out of the first {n} of the workload "
                                      f"(restated CPU path, per-sample throughput; Julia is not installable in this image)"},
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
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e only (no per-kernel hooks, C2 / C5 legs)")
    ap.add_argument("--config", default="c3")
    ap.add_argument("--batch", type=int, default=0,
                    help="override the global minibatch size (profiling the small-minibatch regime of the N-GPU runs on one GPU)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: a C3-sized shard per GPU (default; = config C4 at N=8); strong: config C4 as written, "
                         "8 388 608 transitions in total at any N")
    ap.add_argument("--e2e-feat", default="i8", choices=["i8", "f32"], help="dtype of the host-side features of the e2e leg")
    ap.add_argument("--e2e-mask", default="bits", choices=["bits", "f32"],
                    help="host-side action masks: one bit per action (ppo_buffer_append_packed) or Float32 0 / -Inf")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
