#!/usr/bin/env python
"""Headline benchmark: chamfer fwd+bwd G point-pairs/s at B=16, N=M=4096 (BASELINE.json configs[1], "C2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3|C4]

One "step" = one forward+backward of loss_on_cd (reference model/loss.py:200-205 -> chamfer_dist, loss.py:60-70)
over one synthetic batch (SURVEY.md §8d seeds), through the repo's public API
(`rma_net_b200.model.loss.loss_on_cd` -> autograd.Function -> C ABI -> sm_100a kernels).
Multi-GPU (torchrun, one rank per GPU): batch-sharded, every rank runs its own C2-sized batch, no collective on the
data path ("scaling": "weak"); value = pairs of all ranks / max-over-ranks device time.

The line also carries `secondary`: the other BASELINE.json configurations, measured in the same run (device time, max
over ranks): C3 = the recurrent alignment step (8 x (se3 warp + chamfer), B=16, N=M=8192, every rank its own batch);
C4 = B=512, N=M=2048 split over the ranks by parallel.shard_batch (strong scaling, with the one-GPU time of the whole
batch measured beside it); C5 = one pair N=M=2^20 through parallel.SourceShardedChamfer (sources sharded, NCCL
all-reduce MIN of packed (distance, index) keys and SUM of grad_target inside the timed region) with a self-check
of the merged result against a single-GPU recomputation and the fp64 C oracle on sampled points.

Printed JSON keys beyond the base contract:
  roofline     FP32-FMA roofline of the dominant kernel (chamfer_search_kernel): algorithmic 12 FLOP per point-pair
                / that kernel's mean duration, measured live with CUDA events around the
               `rma_chamfer_search` stage call; peak = SMs x 128 lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json).
  cpu_baseline the reference's chamfer semantics (torch CPU fp32, expanded form + autograd = oracle O1) timed on the
               host cores on a bounded sample, rank 0 at N=1 only.  Reported baseline, not the target.
  e2e          same metric through the public API with HOST (pinned) inputs: every step copies its [B,2N,3] pair batch
               H2D, runs fwd+bwd and reads the loss back (D2H + host sync) inside the timed region.  `value` is the
               MEDIAN of 5 runs of K steps of ONE driver: the double-buffered loop a training step uses (copy of step
               k+1 on a copy stream while step k computes, as rma_net_b200.data.PairBatchLoader does, the host reading
               step k's loss while k+1 runs); `sync_value`: the host reads each loss before launching the next step;
               `serial_value`: nothing overlapped; `copy_in_graph_*`: the same pipeline driven by one graph per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (B, N, M, seed, distribution)  — SURVEY.md §8d
    "C2": (16, 4096, 4096, 0, "uniform"),
    "C3": (16, 8192, 8192, 2, "uniform"),
    "C4": (512, 2048, 2048, 3, "uniform"),
    # small clouds (tools/ only): the size sweep's small rows and the reference's own training shape, T=7 stages x B=4
    # stacked into one batch (train_sample.py:15-30)
    "S16x2048": (16, 2048, 2048, 5, "uniform"),
    "S64x1024": (64, 1024, 1024, 6, "uniform"),
    "S28x2048": (28, 2048, 2048, 7, "uniform"),
    # one rank's shard of C5 at 1 / 2 / 4 / 8 GPUs (tools/ only): 2^20 .. 2^17 sources against 2^20 targets
    "C5s4": (1, 262144, 1048576, 4, "uniform"),
    "C5s8": (1, 131072, 1048576, 4, "uniform"),
    "C5s2": (1, 524288, 1048576, 4, "uniform"),
    "C5s1": (1, 1048576, 1048576, 4, "uniform"),
}
GATE_CYCLES = 40_000_000    # ~20 ms at 1.9 GHz: see the headline leg in run_ours
L2_BYTES = 126 << 20        # B200 L2 (guides/B200_PROFILING.md)
FLOP_PER_PAIR = 12.0        # 6 FMA: two NN searches x 3 multiply-adds
METRIC = "chamfer fwd+bwd G point-pairs/s at B=16,N=M=4096; % of FP32 FMA peak"
UNIT = "G point-pairs/s"


def make_workload(name: str, rank: int = 0):
    B, N, M, seed, _ = WORKLOADS[name]
    g = torch.Generator().manual_seed(seed + 1000 * rank)
    x = torch.rand(B, N, 3, generator=g) * 2 - 1
    y = torch.rand(B, M, 3, generator=g) * 2 - 1
    return x, y


def search_traffic(workload: str, kernel: str):
    """DRAM bytes per launch of the search kernel from the committed ncu capture (profiles/search_dram_traffic.json:
    {"workload", "kernel", "dram_bytes_read", "dram_bytes_write", "source"}), or null when there is none for this
    workload / kernel."""
    p = os.path.join(ROOT, "profiles", "search_dram_traffic.json")
    try:
        with open(p) as f:
            for e in json.load(f):
                if e["workload"] == workload and e["kernel"] == kernel:
                    return {"traffic": float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]),
                            "traffic_source": e["source"]}
    except (OSError, ValueError, KeyError):
        pass
    return {"traffic": None, "traffic_source": None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region (NVML; nvidia-smi as a fallback)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else \
                    self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None,
               "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        return out


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: reference semantics on host cores (oracle O1 = loss.py:45-70 + autograd)
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_time(x, y, sample_b: int, reps: int, threads: int):
    """fwd+bwd of loss_on_cd on the first `sample_b` batch elements; returns best seconds."""
    from oracle.chamfer_oracle import ref_fwd_bwd
    torch.set_num_threads(threads)
    xs, ys = x[:sample_b].contiguous(), y[:sample_b].contiguous()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        ref_fwd_bwd(xs, ys)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    B, N, M, _, _ = WORKLOADS[name]
    x, y = make_workload(name)
    threads = os.cpu_count() or 1
    sample_b = B            # the whole batch: same config as our arm (~0.6 s per step at C2 on 16 cores)
    from oracle.chamfer_oracle import ref_fwd_bwd
    torch.set_num_threads(threads)
    xs, ys = x[:sample_b].contiguous(), y[:sample_b].contiguous()
    for _ in range(args.warmup):
        ref_fwd_bwd(xs, ys)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_fwd_bwd(xs, ys)
    dt = (time.perf_counter() - t0) / args.steps
    pairs = sample_b * N * M
    val = pairs / dt / 1e9
    sample = (f"all {sample_b} of {B} batch elements per step (N=M={N}), torch CPU fp32, {threads} threads: the "
              f"reference formulation (expanded-form [B,N,M] matrix + torch autograd, model/loss.py:45-70, 200-205)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: chamfer fwd+bwd of loss_on_cd, B={B} per GPU, N=M={N}, uniform cube, "
                               f"seed {WORKLOADS[name][3]} (SURVEY.md §8d)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------------------
# secondary configurations of BASELINE.json (C3, C4, C5), measured in the same run
# ------------------------------------------------------------------------------------------------------------
def _timed_max(fn, reps, warm, dev, world):
    """Mean device milliseconds per call of fn over `reps` calls (CUDA events on the launching stream, after `warm`
    untimed calls, barrier + synchronize on both sides), MAX over ranks."""
    import torch.distributed as dist
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _graphed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def secondary_c3(dev, world, rank, peak_fma):
    """BASELINE.json configs[2]: 8 iterations of se3 exp-map warp + chamfer, B=16 (per GPU), N=M=8192, fwd+bwd to
    every phi_k / w_k, gamma = 0.8 (network.py:567-570, 606, 652-655 + loss.py:260-267; SURVEY.md section 8d seeds).
    The network produces all stage outputs, Loss.forward then evaluates the stage loop - here as ONE batched chamfer
    call (loss_cd_stages_fused); replayed as a CUDA graph."""
    from rma_net_b200.model.exfunc.functions import Warp
    from rma_net_b200.model.loss import loss_cd_stages_fused
    B, N, T, gamma = 16, 8192, 8, 0.8
    g = torch.Generator().manual_seed(2 + 1000 * rank)
    src = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    tgt = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    phis = [(torch.randn(B, 6, generator=g) * 0.1).to(dev).requires_grad_(True) for _ in range(T)]
    ws = [torch.rand(B, 1, N, generator=g).to(dev).requires_grad_(True) for _ in range(T)]
    warp = Warp()

    def step():
        for p in phis + ws:
            p.grad = None
        outs, prev = [], None
        for k in range(T):
            out_k, _, _ = warp(phis[k], src, None if k == 0 else ws[k], prev)
            outs.append(out_k.transpose(1, 2))          # [B,3,N] as the network holds the stage outputs
            prev = out_k
        loss = loss_cd_stages_fused(outs, tgt, gamma=gamma)
        loss.backward()
        return loss

    graph = _graphed(step)
    ms = _timed_max(graph.replay, 10, 2, dev, world)
    pairs = T * B * N * N
    return {"what": "8 x (se3 exp-map warp + chamfer) fwd+bwd, B=16 per GPU, N=M=8192, gamma=0.8, CUDA graph; the 8 "
                    "stage chamfers run as one batched call (loss_cd_stages_fused)",
            "n_gpus": world, "scaling": "weak", "ms_per_step": ms, "G_pairs_per_s": world * pairs / ms / 1e6,
            "fp32_fma_peak_frac_per_gpu": 6.0 * pairs / (ms * 1e-3) / peak_fma,
            "gpu_launches_per_step": 2 * T * 2 + 4}


def secondary_c4(dev, world, rank, peak_fma):
    """BASELINE.json configs[3]: B=512, N=M=2048 split contiguously over the ranks (parallel.shard_batch), no
    collective; strong scaling: the one-GPU time of the WHOLE batch is measured beside it on every rank."""
    from rma_net_b200 import parallel
    from rma_net_b200.model.loss import loss_on_cd
    B, N, _, seed, _ = WORKLOADS["C4"]
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, 3, generator=g) * 2 - 1
    y = torch.rand(B, N, 3, generator=g) * 2 - 1
    xd_full, yd_full = x.to(dev), y.to(dev)

    def make(xs_, ys_):
        xr, yr = xs_.detach().requires_grad_(True), ys_.detach().requires_grad_(True)

        def step():
            xr.grad = None
            yr.grad = None
            loss_on_cd(xr, yr).backward()
        return _graphed(step)

    xs_, ys_ = parallel.shard_batch(xd_full, rank, world), parallel.shard_batch(yd_full, rank, world)
    ms = _timed_max(make(xs_, ys_).replay, 20, 3, dev, world)
    pairs = B * N * N
    out = {"what": f"chamfer fwd+bwd, B=512 split over {world} GPU(s) ({xs_.shape[0]} per GPU), N=M=2048, CUDA graph, "
                   f"no collective",
           "n_gpus": world, "scaling": "strong", "ms_per_step": ms, "G_pairs_per_s": pairs / ms / 1e6,
           "fp32_fma_peak_frac_per_gpu": 6.0 * pairs / world / (ms * 1e-3) / peak_fma}
    if world > 1:
        ms1 = _timed_max(make(xd_full, yd_full).replay, 10, 3, dev, world)
        out["one_gpu_ms_per_step"] = ms1
        out["strong_scaling_efficiency"] = ms1 / (world * ms)
    return out


def secondary_c5(dev, world, rank, peak_fma):
    """BASELINE.json configs[4]: one pair N=M=2^20, SOURCE points sharded over the ranks, target replicated
    (parallel.SourceShardedChamfer): the target->source direction is merged by one NCCL all-reduce MIN over packed
    (distance, index) keys, grad_target by one all-reduce SUM; both collectives are inside the timed fwd+bwd.
    Self-check (outside the timed region): the merged dist2/idx2 on 4096 sampled targets and dist1/idx1 on 4096 sampled
    local sources are BIT-equal to a single-GPU recomputation against the full clouds, and 256 of the targets agree
    with the fp64 C oracle (index equal or a near-tie below fp32 resolution, distance within 1e-5)."""
    import torch.distributed as dist
    from rma_net_b200 import ext, parallel
    N = M = 1 << 20
    g = torch.Generator().manual_seed(4)
    x_full = torch.rand(1, N, 3, generator=g) * 2 - 1
    y_full = torch.rand(1, M, 3, generator=g) * 2 - 1
    s0, s1 = parallel.shard_range(N, rank, world)
    xl = x_full[:, s0:s1].to(dev).requires_grad_(True)
    yd = y_full.to(dev).requires_grad_(True)
    keep = {}

    def step():
        xl.grad = None
        yd.grad = None
        d1, d2, i1, i2 = parallel.SourceShardedChamfer.apply(xl, yd, s0, None)
        loss = (d1.sum() + d2.sum()) * 0.5
        loss.backward()
        keep.update(d1=d1.detach(), d2=d2.detach(), i1=i1, i2=i2)

    ms = _timed_max(step, 3, 3, dev, world)      # 3 warm-up steps: NCCL connects lazily, the 9 GiB workspace is cached
    pairs = N * M
    # ---- self-check ------------------------------------------------------------------------------------------
    ok = True
    notes = {}
    gs = torch.Generator().manual_seed(99)
    samp_t = torch.randperm(M, generator=gs)[:4096].to(dev)
    samp_s = torch.randperm(s1 - s0, generator=gs)[:4096].to(dev)
    with torch.no_grad():
        xf = x_full.to(dev)
        _, d2s, _, i2s = ext.chamfer_forward(xf, yd.detach()[:, samp_t].contiguous())     # all sources x sampled targets
        eq_t = bool(torch.equal(d2s[0], keep["d2"][0, samp_t]) and torch.equal(i2s[0], keep["i2"][0, samp_t]))
        d1s, _, i1s, _ = ext.chamfer_forward(xl.detach()[:, samp_s].contiguous(), yd.detach())   # sampled local sources
        eq_s = bool(torch.equal(d1s[0], keep["d1"][0, samp_s]) and torch.equal(i1s[0], keep["i1"][0, samp_s]))
        del xf
    ok &= eq_t and eq_s
    notes["targets_bit_equal_4096"] = eq_t
    notes["local_sources_bit_equal_4096"] = eq_s
    # every rank holds the same merged result
    chk = torch.stack([keep["i2"].sum(dtype=torch.int64), keep["d2"].double().sum().view(torch.int64)])
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    ok &= same
    notes["ranks_hold_identical_merge"] = same
    if rank == 0:
        try:
            import numpy as np
            from oracle import c_oracle           # the checker (test infrastructure), never the thing measured
            t256 = samp_t[:256].cpu()
            e = c_oracle.chamfer_fwd(x_full.numpy(), y_full[:, t256].numpy(), "f64")
            d_gpu = keep["d2"][0, samp_t[:256]].double().cpu().numpy()
            i_gpu = keep["i2"][0, samp_t[:256]].cpu().numpy()
            rel = np.abs(d_gpu - e["dist2"][0]) / np.maximum(e["dist2"][0], 1e-30)
            xf64 = x_full[0].double().numpy()
            yt = y_full[0, t256].double().numpy()
            d_at_gpu_idx = ((xf64[i_gpu] - yt) ** 2).sum(-1)
            near_tie = np.abs(d_at_gpu_idx - e["dist2"][0]) <= 2.0 ** -20 * e["dist2"][0]
            oracle_ok = bool((rel <= 1e-5).all() and ((i_gpu == e["idx2"][0]) | near_tie).all())
            notes["fp64_oracle_256_targets"] = oracle_ok
            notes["fp64_oracle_max_rel_err"] = float(rel.max())
            notes["fp64_oracle_index_mismatches_all_near_ties"] = int((i_gpu != e["idx2"][0]).sum())
            ok &= oracle_ok
        except Exception as ex:       # the oracle library is missing on this box: say so, do not pass silently
            notes["fp64_oracle_256_targets"] = f"not run: {type(ex).__name__}: {ex}"
            ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    path = ext.chamfer_path(1, s1 - s0, M)
    return {"what": f"one pair N=M=2^20, sources sharded over {world} GPU(s) ({s1 - s0} per GPU), target replicated; "
                    f"fwd+bwd incl. NCCL all-reduce MIN of packed (distance,index) keys [M] and SUM of grad_target",
            "n_gpus": world, "scaling": "strong", "ms_per_step": ms, "G_pairs_per_s": pairs / ms / 1e6,
            "fp32_fma_peak_frac_per_gpu": 6.0 * pairs / world / (ms * 1e-3) / peak_fma,
            "search_path": path, "collectives_per_step": 0 if world == 1 else 2,
            "c5_parity": bool(flag.item() == 1), "c5_checks": notes}

# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from rma_net_b200 import _lib, ext
    from rma_net_b200.model.loss import loss_on_cd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    name = args.workload
    B, N, M, _, _ = WORKLOADS[name]
    pairs = B * N * M
    x_host, y_host = make_workload(name, rank)
    x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()
    xd = x_host.to(dev).requires_grad_(True)
    yd = y_host.to(dev).requires_grad_(True)

    def step():
        xd.grad = None
        yd.grad = None
        loss = loss_on_cd(xd, yd)
        loss.backward()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- capture the step in a CUDA graph (launch-bound otherwise: ~10 small launches per step) --------------
    use_graph = not args.no_graph
    graph = None
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = step()
        run = graph.replay
    else:
        run = step

    # ---- headline leg: K steps back to back over ROTATING input sets larger than L2 ----------------------------
    # A step's inputs are 1.5 MiB, L2-resident by nature if the same pair were replayed.  So the timed steps walk
    # through R distinct device-resident input sets (own clouds, own gradient tensors, own CUDA graph) whose total size
    # is > 2 x L2: every timed step reads inputs that are not in L2 and writes gradients that are not, and there is no
    # host work, event or flush kernel between the steps - the way a training loop issues them.  The graphs share ONE
    # memory pool, so the library's scratch (workspace, gradient halves: freed at the end of every step) is the same
    # block for every set, as the caching allocator hands a training loop the same block every step; --private-pools
    # gives every set its own scratch as well (round-2 protocol until r02k: ~3 us per step of record lines allocating
    # in L2 behind the previous steps' dirty ones).  One event pair around exactly K steps.
    W = max(args.warmup, 3)
    set_bytes = 4 * (xd.numel() + yd.numel()) * 2               # clouds + their gradients
    K = args.steps
    # Steps per CUDA graph: the K steps of a timed region are replayed as K / n graphs of n consecutive steps on n
    # DIFFERENT input sets (n = K up to 40, else the largest divisor of K up to 40).  One graph per step was the protocol
    # until r02k; it reads ~4 us per step more for the same kernels on the same cold inputs, which is what switching
    # between 93 different graph executables costs the launch path (tools/rotate_probe.py: one graph replayed 72.0 us,
    # 93 graphs on ONE pair 76.1, 93 graphs on 93 pairs 77.0, graphs of 31 steps on 31 pairs 70.6) - a training loop
    # replays one executable.  That figure is still measured and reported as `one_graph_per_step`.
    n_per = 1
    if use_graph and not args.one_graph_per_step:
        cap = args.steps_per_graph if args.steps_per_graph > 0 else 40
        n_per = max(d for d in range(1, min(K, cap) + 1) if K % d == 0)
    R = args.sets if args.sets > 0 else max(2, int(2.2 * L2_BYTES // set_bytes) + 1)
    R = ((R + n_per - 1) // n_per) * n_per
    if R // n_per < 2:
        R = 2 * n_per
    step_fns, keep_alive = [], []
    shared_pool = not (args.private_pools or not use_graph) or None   # reporting: do the graphs share scratch?
    gen = torch.Generator().manual_seed(1234 + rank)
    for r in range(R):
        if r == 0:
            xr_, yr_ = xd, yd
        else:       # the same distribution as the workload (uniform cube), different points
            xr_ = (torch.rand(B, N, 3, generator=gen) * 2 - 1).to(dev).requires_grad_(True)
            yr_ = (torch.rand(B, M, 3, generator=gen) * 2 - 1).to(dev).requires_grad_(True)

        def step_r(xr_=xr_, yr_=yr_):
            xr_.grad = None
            yr_.grad = None
            loss_r = loss_on_cd(xr_, yr_)
            loss_r.backward()
            return loss_r
        keep_alive.append((xr_, yr_))          # a graph holds raw pointers: its inputs must outlive the closure
        for _ in range(2):
            step_r()
        step_fns.append(step_r)

    def capture_groups(per):
        """R / per callables, each running `per` consecutive steps (sets i*per .. i*per+per-1)."""
        if not use_graph:
            return [lambda i=i: [f() for f in step_fns[i:i + per]] for i in range(0, R, per)]
        out = []
        shared_pool = None if args.private_pools else torch.cuda.graph_pool_handle()   # one pool per family of graphs
        for i in range(0, R, per):
            torch.cuda.synchronize()
            g_r = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_r, pool=shared_pool):
                for f in step_fns[i:i + per]:
                    f()
            out.append(g_r.replay)
        return out

    def timed_regions(groups, per, nrep):
        """nrep regions of exactly K steps (K / per launches each), one event pair per region; ms per step."""
        nlaunch = K // per
        barrier()
        for g_ in groups:         # the first launch of a graph uploads it: replay every group once, in order, untimed
            g_()                  # (R >= W warm-up steps on inputs that are then > 2 x L2 of traffic old)
        barrier()
        out = []
        pos = 0
        for rep in range(nrep):
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # A ~20 ms spin kernel in front of the start event: the host enqueues all launches while it runs, so the K
            # steps execute back to back on the device whatever the host does meanwhile (the clock sampler's NVML
            # queries stall launches for milliseconds: without the gate a region now and then measured such a stall)
            torch.cuda._sleep(GATE_CYCLES)
            for _ in range(-(-W // per)):     # >= W warm-up steps of the same kind directly in front of the start
                groups[pos % len(groups)]()   # event (other input sets): the first steps behind the idle spin ran
                pos += 1                      # ~1.5 us per step slower at K = 20
            ev0.record()
            for _ in range(nlaunch):
                groups[pos % len(groups)]()
                pos += 1
            ev1.record()
            barrier()
            out.append(ev0.elapsed_time(ev1) / K)
        return out

    sets = capture_groups(n_per)
    sampler = ClockSampler(local)
    sampler.start()
    region_ms = timed_regions(sets, n_per, 5)   # region 0 is THE measurement (exactly K steps); the others give its spread
    ms = region_ms[0]
    if n_per > 1:                 # the r02a-r02k protocol on the same inputs: one graph (executable) per step
        del sets
        one_per_ms = timed_regions(capture_groups(1), 1, 3)
    else:
        one_per_ms = list(region_ms[:3])

    # ---- flushed leg (the round-1 protocol, kept for comparison): one input set, a 256 MiB memset between the steps,
    # one event pair per step.  It adds the idle gap between an event record and the first kernel of a graph launch
    # and the write-back of the memset's dirty lines to every step.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    for _ in range(W):
        flush.zero_()
        run()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                 # L2 flush between timed iterations (outside the per-step event bracket)
        starts[k].record()
        run()
        ends[k].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [s_.elapsed_time(e_) for s_, e_ in zip(starts, ends)]
    flushed_ms = sum(step_ms) / len(step_ms)

    # ---- roofline leg: the dominant kernel alone, events around the rma_chamfer_search stage -------------------
    lib = _lib.load()
    import ctypes
    wsb = lib.rma_chamfer_workspace_bytes(B, N, M, None)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    d1 = torch.empty(B, N, device=dev)
    d2 = torch.empty(B, M, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
    i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    xs_, ys_ = xd.stride(), yd.stride()
    k_ms = []
    for it in range(args.steps + 3):
        flush.zero_()
        _lib.check(lib.rma_chamfer_prep(P(xd), xs_[0], xs_[1], xs_[2], P(yd), ys_[0], ys_[1], ys_[2], B, N, M,
                                        P(ws), wsb, None, stream), "prep")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rma_chamfer_search(B, N, M, P(ws), wsb, None, stream), "search")
        e1.record()
        _lib.check(lib.rma_chamfer_finalize(B, N, M, P(d1), P(d2), P(i1), P(i2), P(ws), wsb, None, stream),
                   "finalize")
        torch.cuda.synchronize()
        if it >= 3:
            k_ms.append(e0.elapsed_time(e1))
    search_ms = sum(k_ms) / len(k_ms)
    # the same kernel, 8 launches back to back between one event pair (no event / launch gap per launch, records
    # overwritten in place): what a launch costs inside a step, where the device timeline has it at 54.3 us
    # (profiles/r02n_step_timeline_C2.json).  Reported beside `frac`, which stays the flushed single launch.
    kb_ms = []
    for it in range(4):
        flush.zero_()
        _lib.check(lib.rma_chamfer_prep(P(xd), xs_[0], xs_[1], xs_[2], P(yd), ys_[0], ys_[1], ys_[2], B, N, M,
                                        P(ws), wsb, None, stream), "prep")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            _lib.check(lib.rma_chamfer_search(B, N, M, P(ws), wsb, None, stream), "search")
        e1.record()
        _lib.check(lib.rma_chamfer_finalize(B, N, M, P(d1), P(d2), P(i1), P(i2), P(ws), wsb, None, stream),
                   "finalize")
        torch.cuda.synchronize()
        if it >= 1:
            kb_ms.append(e0.elapsed_time(e1) / 8)
    search_b2b_ms = sum(kb_ms) / len(kb_ms)
    clocks = sampler.stop()
    # which search implementation ran for this size, and how often the filter's guard sent an item to the slow path
    stats = (ctypes.c_uint * 4)()
    filt = 1 if lib.rma_chamfer_path(B, N, M, None) == _lib.PATH_FILTER else 0
    _lib.check(lib.rma_chamfer_guard_stats(B, N, M, P(ws), None, ctypes.cast(stats, ctypes.c_void_p), stream),
               "guard_stats")
    search_kernel = "chamfer_fsearch_kernel" if filt == 1 else "chamfer_search_kernel"

    # ---- e2e leg: host (pinned) inputs, public API, H2D + fwd+bwd + D2H loss every step -------------------------
    # The host batch is laid out the way the reference holds it (train_view.py:89-100, 369-370): one [B, 2N, 3] pair
    # batch whose halves are source and target.  One H2D copy brings both clouds over; the op takes the two strided
    # slices (batch stride 6N) as they are.
    loss_pin = torch.zeros((), dtype=torch.float32).pin_memory()
    pair_pin = torch.cat([x_host, y_host], 1).pin_memory() if N == M else None
    if pair_pin is not None:
        pair_d = torch.empty_like(pair_pin, device=dev)
        xe = pair_d[:, :N].detach().requires_grad_(True)
        ye = pair_d[:, N:].detach().requires_grad_(True)
        h2d_bytes = pair_pin.numel() * 4
    else:
        xe, ye = xd, yd
        h2d_bytes = x_pin.numel() * 4 + y_pin.numel() * 4

    def e2e_compute():
        xe.grad = None
        ye.grad = None
        loss = loss_on_cd(xe, ye)
        loss.backward()
        return loss

    def e2e_body():
        with torch.no_grad():
            if pair_pin is not None:
                pair_d.copy_(pair_pin, non_blocking=True)
            else:
                xd.copy_(x_pin, non_blocking=True)
                yd.copy_(y_pin, non_blocking=True)
        loss = e2e_compute()
        loss_pin.copy_(loss.detach(), non_blocking=True)
        return loss

    if use_graph:
        # the same public-API step, with the H2D copy and the D2H copy of the loss as graph nodes
        for _ in range(3):
            e2e_body()
        torch.cuda.synchronize()
        e2e_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e2e_graph):
            e2e_body()

        def e2e_step():
            e2e_graph.replay()
            torch.cuda.current_stream().synchronize()     # the host reads the loss every step
            return float(loss_pin)
    else:
        def e2e_step():
            e2e_body()
            torch.cuda.current_stream().synchronize()
            return float(loss_pin)

    # Every e2e region below: barrier, then W warm-up steps of the same loop, then - without another synchronisation in
    # between - the start event and exactly K timed steps.  (With the barrier directly in front of the start event the
    # first step of a region started on an idle GPU; on some boxes that cost ~1.2 ms per region, i.e. 60 us per step at
    # K = 20 and 6 us at K = 200.)
    E2E_W = max(args.warmup, 3)
    barrier()
    for _ in range(E2E_W):
        e2e_step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        serial_loss = e2e_step()
    ev1.record()
    barrier()
    e2e_serial_ms = ev0.elapsed_time(ev1) / args.steps
    e2e_ms, e2e_mode = e2e_serial_ms, "serial: H2D copy, fwd+bwd, D2H loss, host sync, one after the other"
    e2e_sync_ms = e2e_serial_ms
    e2e_extra = {}

    if use_graph and pair_pin is not None:
        # The same K steps the way a training loop feeds them (rma_net_b200.data.PairBatchLoader, the replacement of
        # the reference's per-step `.cuda()`, train_view.py:359-366): two device buffers, the H2D copy of step k+1 on
        # a copy stream while step k computes.  Every step still copies its own 1.5 MB from pinned host memory and
        # the host still reads every step's loss before it launches the next step; K steps = K copies + K reads.
        main_stream = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        bufs = [torch.empty_like(pair_pin, device=dev) for _ in range(2)]
        xs = [b_[:, :N].detach().requires_grad_(True) for b_ in bufs]
        ys = [b_[:, N:].detach().requires_grad_(True) for b_ in bufs]

        loss_pins = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

        d2h_stream = torch.cuda.Stream()

        def compute(i):
            # the 4-byte D2H copy of the loss is issued as soon as the forward has produced it, on a side stream (a
            # parallel branch of the captured graph), so that the copy engine's latency runs beside the backward
            # kernels instead of behind them
            xs[i].grad = None
            ys[i].grad = None
            loss = loss_on_cd(xs[i], ys[i])
            cur = torch.cuda.current_stream()
            d2h_stream.wait_stream(cur)
            with torch.cuda.stream(d2h_stream):
                loss_pins[i].copy_(loss.detach(), non_blocking=True)
            loss.backward()
            cur.wait_stream(d2h_stream)

        graphs = []
        for i in range(2):
            bufs[i].copy_(pair_pin)
            for _ in range(3):
                compute(i)
            torch.cuda.synchronize()
            g_i = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_i):
                compute(i)
            graphs.append(g_i)
        ev_copied = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]

        def issue_copy(i):
            copy_stream.wait_event(ev_free[i])               # the step that last used this buffer has finished
            with torch.cuda.stream(copy_stream):
                bufs[i].copy_(pair_pin, non_blocking=True)
            ev_copied[i].record(copy_stream)

        def run_pipelined(steps, first_event=None, lag=0):
            """lag=0: the host reads step k's loss before it launches step k+1.  lag=1: it launches step k+1 first and
            reads step k's loss (own pinned slot, own event) while k+1 runs - the logging pattern of a training loop
            that never stalls the device; still K copies in, K losses read, all inside the timed region."""
            for i in range(2):
                ev_free[i].record(main_stream)
            if first_event is not None:
                copy_stream.wait_event(first_event)          # the first copy starts inside the timed region
            issue_copy(0)
            vals = []
            for k in range(steps):
                i = k & 1
                main_stream.wait_event(ev_copied[i])
                graphs[i].replay()
                ev_free[i].record(main_stream)               # also marks "loss of step k is in its pinned slot"
                if k + 1 < steps:
                    issue_copy(1 - i)
                if lag == 0:
                    main_stream.synchronize()                # the host reads the loss every step
                    vals.append(float(loss_pins[i]))
                elif k >= 1:
                    ev_free[1 - i].synchronize()             # step k-1 done (step k is running)
                    vals.append(float(loss_pins[1 - i]))
            if lag:
                main_stream.synchronize()
                vals.append(float(loss_pins[(steps - 1) & 1]))
            assert len(vals) == steps
            return vals[-1]

        # Every pipelined e2e figure below is the MEDIAN of E2E_REPS runs of K steps each (the host is inside every
        # one of these loops; min and max of the repeats are reported beside the headline driver).
        E2E_REPS = 5
        e2e_modes, e2e_spread = {}, {}
        for lag in (0, 1):
            run_pipelined(4, lag=lag)
            reps_ms = []
            for _ in range(E2E_REPS):
                barrier()
                run_pipelined(E2E_W, lag=lag)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                pipelined_loss = run_pipelined(args.steps, ev0, lag=lag)
                ev1.record()
                barrier()
                if abs(pipelined_loss - serial_loss) > 1e-5 * abs(serial_loss):
                    raise RuntimeError(f"pipelined e2e loss {pipelined_loss} != serial {serial_loss}")
                reps_ms.append(ev0.elapsed_time(ev1) / args.steps)
            e2e_modes[lag] = statistics.median(reps_ms)
            e2e_spread[lag] = [min(reps_ms), max(reps_ms)]
        e2e_sync_ms = e2e_modes[0]
        e2e_ms = e2e_modes[1]

        # Same pipeline with the copy of step k+1 as a parallel branch INSIDE step k's graph (fork / join on the copy
        # stream during capture): one graph launch, one event record and one event wait per step on the host.  On a
        # box with slow host cores the Python-driven event choreography above costs more than the device step.
        fused = []
        for i in range(2):
            g_i = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_i):
                cap = torch.cuda.current_stream()
                copy_stream.wait_stream(cap)
                with torch.cuda.stream(copy_stream):
                    bufs[1 - i].copy_(pair_pin, non_blocking=True)        # inputs of step k+1
                compute(i)                                                # step k on bufs[i]
                cap.wait_stream(copy_stream)
            fused.append(g_i)
        ev_done = [torch.cuda.Event() for _ in range(2)]

        def run_fused(steps):
            bufs[0].copy_(pair_pin, non_blocking=True)                    # inputs of step 0 (inside the timed region)
            vals = []
            for k in range(steps):
                i = k & 1
                fused[i].replay()
                ev_done[i].record(main_stream)
                if k >= 1:
                    ev_done[1 - i].synchronize()
                    vals.append(float(loss_pins[1 - i]))
            main_stream.synchronize()
            vals.append(float(loss_pins[(steps - 1) & 1]))
            assert len(vals) == steps
            return vals[-1]

        run_fused(4)
        fused_ms = []
        for _ in range(E2E_REPS):
            barrier()
            run_fused(E2E_W)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            fused_loss = run_fused(args.steps)
            ev1.record()
            barrier()
            if abs(fused_loss - serial_loss) > 1e-5 * abs(serial_loss):
                raise RuntimeError(f"graph-pipelined e2e loss {fused_loss} != serial {serial_loss}")
            fused_ms.append(ev0.elapsed_time(ev1) / args.steps)
        e2e_fused_ms = statistics.median(fused_ms)
        # THE e2e figure is the copy_in_graph driver's (one graph launch, one event record and one event wait per step on
        # the host): on boxes with slow host cores the Python event choreography of the host loop is what its figure
        # measures (r02n box: 114 us host loop, 95 us in-graph, 229 us serial; a faster host: 81.5 / 81.6 / 122).  Both
        # are medians of E2E_REPS runs of K steps, both are reported.
        host_loop_ms = e2e_ms
        e2e_ms = e2e_fused_ms
        e2e_extra = {"driver": "copy_in_graph", "repeats": E2E_REPS,
                     "ms_per_step_min_max": [min(fused_ms), max(fused_ms)], "sync_ms_per_step_min_max": e2e_spread[0],
                     "copy_stream_host_loop_ms_per_step": host_loop_ms,
                     "copy_stream_host_loop_value": world * pairs / (host_loop_ms * 1e-3) / 1e9,
                     "copy_stream_host_loop_ms_per_step_min_max": e2e_spread[1],
                     "note": "value = MEDIAN of 5 runs of K steps of the copy_in_graph driver: the H2D copy of step k+1 "
                             "is a parallel branch of step k's CUDA graph (second device buffer), the 4-byte D2H copy of "
                             "the loss a side branch beside the backward; per step the host launches one graph, records "
                             "one event and waits for the previous step's event before it reads that step's loss from "
                             "its pinned slot.  copy_stream_host_loop_*: the same pipeline with the copy issued from "
                             "Python on a copy stream with events, median of 5 as well; reported beside it, not folded "
                             "into value"}
        e2e_mode = ("double-buffered + lagged read: the H2D copy of step k+1 overlaps step k (second device buffer; "
                    "rma_net_b200.data.PairBatchLoader pattern) and the host reads step k's loss (own pinned "
                    "slot, event) while step k+1 runs; every step copies its inputs and every loss is read by the host "
                    "inside the timed region.  sync_ms_per_step: the host-loop driver with the host reading each loss "
                    "before it launches the next step; serial_ms_per_step: nothing overlapped.  Pipelined figures: median "
                    "of 5 runs of K steps each")

    # ---- secondary configurations (C3, C4, C5) -----------------------------------------------------------------
    info = ext.device_info()
    peaks, peak_src = measured_peaks()
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_fma = info["sm_count"] * 128 * sm_max_mhz * 1e6          # FMA/s per GPU
    secondary = {}
    if not args.no_secondary:
        del flush
        torch.cuda.empty_cache()
        for key, fn in (("C3", secondary_c3), ("C4", secondary_c4), ("C5", secondary_c5)):
            try:
                secondary[key] = fn(dev, world, rank, peak_fma)
            except Exception as ex:          # a failed secondary must not take the headline line with it
                secondary[key] = {"error": f"{type(ex).__name__}: {ex}"}
            torch.cuda.empty_cache()

    # ---- aggregate over ranks (max time) ------------------------------------------------------------------------
    t = torch.tensor([ms, e2e_ms, search_ms, e2e_serial_ms, e2e_sync_ms, flushed_ms] + region_ms + one_per_ms,
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, search_ms, e2e_serial_ms, e2e_sync_ms, flushed_ms = (float(v) for v in t.cpu()[:6])
    region_ms = [float(v) for v in t.cpu()[6:6 + len(region_ms)]]
    one_per_ms = [float(v) for v in t.cpu()[6 + len(region_ms):]]
    value = world * pairs / (ms * 1e-3) / 1e9
    e2e_value = world * pairs / (e2e_ms * 1e-3) / 1e9

    peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    achieved_tflops = FLOP_PER_PAIR * pairs / (search_ms * 1e-3) / 1e12
    ctas, thr, spl, occ = (ctypes.c_int() for _ in range(4))
    lib.rma_chamfer_search_config(B, N, M, None, ctypes.byref(ctas), ctypes.byref(thr), ctypes.byref(spl),
                                  ctypes.byref(occ))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: chamfer fwd+bwd of loss_on_cd, B={B} per GPU, N=M={N}, uniform cube, "
                               f"seed {WORKLOADS[name][3]} (SURVEY.md §8d)",
                   "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": f"inputs larger than L2: the timed steps rotate through {R} device-resident input sets "
                         f"({R * set_bytes / 2**20:.0f} MiB of clouds + gradients > 2 x 126 MiB L2), replayed back to back "
                         f"as {args.steps // n_per} CUDA graph(s) of {n_per} consecutive steps on {n_per} different sets, "
                         f"one event pair around the K steps; "
                         + ("every set's graph has its own memory pool (own scratch workspace too); " if shared_pool is None
                            else "the graphs share one memory pool, so the library's scratch workspace is the same block "
                                 "every step (as an allocator gives a training loop), inputs and gradients are not; ")
                         + "flushed_step = the round-1 protocol (one set, 256 MiB memset + one event pair per step) for "
                           "comparison",
                   "input_sets": R,
                   "steps_per_graph": n_per,
                   "scratch": "private per set" if shared_pool is None else "shared pool",
                   "cuda_graph": bool(use_graph),
                   "search_launch": {"kernel": search_kernel, "ctas": ctas.value, "threads": thr.value,
                                     ("units32_per_cta" if filt == 1 else "column_splits"): spl.value,
                                     "ctas_per_sm": occ.value},
                   "search_path": ("filter (expanded form, 3 FFMA + 1 FADD per pair) + exact difference-form resolve"
                                   if filt == 1 else "exact difference form (6 FP32 ops per pair)"),
                   "guard_slow_path": ({"rows": int(stats[0]), "columns": int(stats[1]),
                                        "row_block_rescans": int(stats[2]), "of_items": B * (N + M)}
                                       if filt == 1 else None)},
        "fp32_fma_peak_frac_step": FLOP_PER_PAIR * pairs / (ms * 1e-3) / 1e12 / peak_tflops,
        "roofline": {"bound": "fp32_fma", "kernel": search_kernel, "achieved": achieved_tflops,
                     "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
                     # the north-star quantity: the same algorithmic FLOP over the WHOLE fwd+bwd step (prep + search +
                     # resolve + backward) of the headline leg (inputs rotating through > 2 x L2)
                     "step_frac": FLOP_PER_PAIR * pairs / (ms * 1e-3) / 1e12 / peak_tflops,
                     "step_frac_flushed_protocol": FLOP_PER_PAIR * pairs / (flushed_ms * 1e-3) / 1e12 / peak_tflops,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel per launch cannot be measured inside
                     # this run (it needs ncu): taken from the committed ncu --set full capture named in
                     # traffic_source, for this workload and search kernel only
                     **search_traffic(name, search_kernel), "kernel_ms": search_ms,
                     "kernel_ms_back_to_back": search_b2b_ms,
                     "frac_back_to_back": FLOP_PER_PAIR * pairs / (search_b2b_ms * 1e-3) / 1e12 / peak_tflops,
                     "note": ("achieved = ALGORITHMIC 12 FLOP (6 FMA) per point-pair (SURVEY.md section 8d) / kernel "
                              "time.  The filter kernel itself issues 3 FFMA + 1 FADD + 2 min per pair (its exact "
                              "re-evaluation of the candidates runs in the resolve kernel), so frac is a fraction "
                              "of the FMA-issue roofline of the reference formulation, not FMA-pipe utilisation; "
                              "ncu pipe numbers are in profiles/." if filt == 1 else
                              "achieved = algorithmic 12 FLOP per point-pair / kernel time"),
                     "peak_source": f"{info['sm_count']} SMs x 128 FP32 lanes x 2 x sm_max_mhz={sm_max_mhz:.0f} "
                                    f"({peak_src} MEASURED_PEAKS.json clock); FFMA microbenchmark on this pool "
                                    f"reaches 97% of it (profiles/r01_fp32_microbench.jsonl)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4,
                "mode": e2e_mode, "sync_ms_per_step": e2e_sync_ms,
                "sync_value": world * pairs / (e2e_sync_ms * 1e-3) / 1e9, "serial_ms_per_step": e2e_serial_ms,
                "serial_value": world * pairs / (e2e_serial_ms * 1e-3) / 1e9,
                "host_layout": "one pinned [B,2N,3] pair batch (train_view.py:369-370), one H2D copy, strided slices"
                               if pair_pin is not None else "two pinned clouds, two H2D copies", **e2e_extra},
        "gpu_launches": 4 * args.steps,      # prep + search + resolve/finalize(+loss,+grads) + cd_backward per step
        "secondary": secondary,
        "clocks": clocks,
        "region_ms_per_step": region_ms,      # [0] is ms_per_step; four more regions of K steps show the spread
        "one_graph_per_step": {"ms_per_step": one_per_ms[0], "region_ms_per_step": one_per_ms,
                               "value": world * pairs / (one_per_ms[0] * 1e-3) / 1e9,
                               "note": "the same K steps on the same rotating input sets with one CUDA graph (executable) "
                                       "per step and set - the headline protocol of r02a-r02k; the difference to "
                                       "ms_per_step is the launch path switching between executables, "
                                       "tools/rotate_probe.py"},
        "flushed_step": {"ms_per_step": flushed_ms, "step_ms_min_max": [min(step_ms), max(step_ms)],
                         "value": world * pairs / (flushed_ms * 1e-3) / 1e9,
                         "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3},
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample_b = B
        sec = cpu_reference_time(x_host, y_host, sample_b, reps=3, threads=threads)
        line["cpu_baseline"] = {
            "value": sample_b * N * M / sec / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"all {sample_b} of {B} batch elements, best of 3, torch CPU fp32 (reference formulation: "
                      f"expanded-form [B,N,M] matrix + autograd, model/loss.py:45-70, 200-205)"}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        # the JSON line must be the LAST line of stdout: NCCL's version banner sits in the C stdio buffer when stdout
        # is a pipe and would otherwise be flushed behind it at exit
        import ctypes as _ct
        sys.stdout.flush()
        _ct.CDLL(None).fflush(None)
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4"])
    ap.add_argument("--sets", type=int, default=0, help="rotating input sets of the headline leg (default: enough for "
                    "2.2 x L2; a small number only for profiler runs, whose line is not a bench value)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--one-graph-per-step", action="store_true", help="headline leg: one CUDA graph per step and input set "
                    "(the protocol until r02k) instead of graphs of n consecutive steps")
    ap.add_argument("--steps-per-graph", type=int, default=0, help="headline leg: cap on the steps captured per CUDA graph "
                    "(default 40; the largest divisor of K below the cap is used)")
    ap.add_argument("--private-pools", action="store_true", help="headline leg: every input set's graph gets its own memory "
                    "pool (its own scratch workspace) instead of the shared one")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C3 / C4 / C5 legs of the `secondary` object")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
