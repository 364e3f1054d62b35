#!/usr/bin/env python
"""Headline benchmark: chamfer fwd+bwd G point-pairs/s at B=16, N=M=4096 (BASELINE.json configs[1], "C2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3|C4]

One "step" = one forward+backward of loss_on_cd (reference model/loss.py:200-205 -> chamfer_dist, loss.py:60-70)
over one synthetic batch (SURVEY.md §8d seeds), through the repo's public API
(`rma_net_b200.model.loss.loss_on_cd` -> autograd.Function -> C ABI -> sm_100a kernels).
Multi-GPU (torchrun, one rank per GPU): batch-sharded, every rank runs its own C2-sized batch, no collective on the
data path ("scaling": "weak"); value = pairs of all ranks / max-over-ranks device time.

Printed JSON keys beyond the base contract:
  roofline     FP32-FMA roofline of the dominant kernel (chamfer_search_kernel): algorithmic 12 FLOP per point-pair
                / that kernel's mean duration, measured live with CUDA events around the
               `rma_chamfer_search` stage call; peak = SMs x 128 lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json).
  cpu_baseline the reference's chamfer semantics (torch CPU fp32, expanded form + autograd = oracle O1) timed on the
               host cores on a bounded sample, rank 0 at N=1 only.  Reported baseline, not the target.
  e2e          same metric through the public API with HOST (pinned) inputs: every step copies its [B,2N,3] pair batch
               H2D, runs fwd+bwd and reads the loss back (D2H + host sync) inside the timed region.  `value` is the
               double-buffered loop a training step uses (copy of step k+1 on a copy stream while step k computes,
               as rma_net_b200.data.PairBatchLoader does); `serial_value` is the same with nothing overlapped.
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
}
FLOP_PER_PAIR = 12.0        # 6 FMA: two NN searches x 3 multiply-adds
METRIC = "chamfer fwd+bwd G point-pairs/s at B=16,N=M=4096; % of FP32 FMA peak"
UNIT = "G point-pairs/s"


def make_workload(name: str, rank: int = 0):
    B, N, M, seed, _ = WORKLOADS[name]
    g = torch.Generator().manual_seed(seed + 1000 * rank)
    x = torch.rand(B, N, 3, generator=g) * 2 - 1
    y = torch.rand(B, M, 3, generator=g) * 2 - 1
    return x, y


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
    sample_b = min(B, 2)
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
    sample = f"first {sample_b} of {B} batch elements per step (N=M={N}), torch CPU fp32, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: chamfer fwd+bwd B={B} N=M={N} (SURVEY.md §8d), bounded CPU sample",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

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

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        run()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                 # L2 flush between timed iterations (outside the per-step event bracket)
        starts[k].record()
        run()
        ends[k].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    ms = sum(step_ms) / len(step_ms)

    # ---- roofline leg: the dominant kernel alone, events around the rma_chamfer_search stage -------------------
    lib = _lib.load()
    import ctypes
    wsb = lib.rma_chamfer_workspace_bytes(B, N, M)
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
                                        P(ws), wsb, stream), "prep")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rma_chamfer_search(B, N, M, P(ws), wsb, stream), "search")
        e1.record()
        _lib.check(lib.rma_chamfer_finalize(B, N, M, P(d1), P(d2), P(i1), P(i2), P(ws), wsb, stream), "finalize")
        torch.cuda.synchronize()
        if it >= 3:
            k_ms.append(e0.elapsed_time(e1))
    search_ms = sum(k_ms) / len(k_ms)
    clocks = sampler.stop()
    # which search implementation ran for this size, and how often the filter's guard sent an item to the slow path
    stats = (ctypes.c_uint * 4)()
    filt = lib.rma_exp_chamfer_path_info(B, N, M, P(ws), ctypes.cast(stats, ctypes.c_void_p), stream)
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

    for _ in range(3):
        e2e_step()
    barrier()
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

        def compute(i):
            xs[i].grad = None
            ys[i].grad = None
            loss = loss_on_cd(xs[i], ys[i])
            loss.backward()
            loss_pins[i].copy_(loss.detach(), non_blocking=True)

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

        # Every e2e figure below is the best of E2E_REPS runs of K steps each: a single host hiccup (a few ms on a
        # shared box) is a quarter of a K = 200 run of 18 ms, and the host is inside every one of these loops.
        E2E_REPS = 3
        e2e_modes = {}
        for lag in (0, 1):
            run_pipelined(4, lag=lag)
            best = float("inf")
            for _ in range(E2E_REPS):
                barrier()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                pipelined_loss = run_pipelined(args.steps, ev0, lag=lag)
                ev1.record()
                barrier()
                if abs(pipelined_loss - serial_loss) > 1e-6 * abs(serial_loss):
                    raise RuntimeError(f"pipelined e2e loss {pipelined_loss} != serial {serial_loss}")
                best = min(best, ev0.elapsed_time(ev1) / args.steps)
            e2e_modes[lag] = best
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
        e2e_fused_ms = float("inf")
        for _ in range(E2E_REPS):
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            fused_loss = run_fused(args.steps)
            ev1.record()
            barrier()
            if abs(fused_loss - serial_loss) > 1e-6 * abs(serial_loss):
                raise RuntimeError(f"graph-pipelined e2e loss {fused_loss} != serial {serial_loss}")
            e2e_fused_ms = min(e2e_fused_ms, ev0.elapsed_time(ev1) / args.steps)
        e2e_extra = {"copy_stream_host_loop_ms_per_step": e2e_ms, "copy_in_graph_ms_per_step": e2e_fused_ms,
                     "note": "value = the faster of two drivers of the same pipeline: the H2D copy of step k+1 "
                             "issued from Python on a copy stream with events, or captured as a parallel branch of "
                             "step k's CUDA graph (fewer host calls per step)"}
        e2e_ms = min(e2e_ms, e2e_fused_ms)
        e2e_mode = ("double-buffered + lagged read: the H2D copy of step k+1 overlaps step k (second device buffer, copy "
                    "stream; rma_net_b200.data.PairBatchLoader pattern) and the host reads step k's loss (own pinned "
                    "slot, event) while step k+1 runs; every step copies its inputs and every loss is read by the host "
                    "inside the timed region.  sync_ms_per_step: same, but the host reads each loss before it launches "
                    "the next step; serial_ms_per_step: nothing overlapped.  Pipelined figures: best of 3 runs of K "
                    "steps each")

    # ---- aggregate over ranks (max time) ------------------------------------------------------------------------
    t = torch.tensor([ms, e2e_ms, search_ms, e2e_serial_ms, e2e_sync_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, search_ms, e2e_serial_ms, e2e_sync_ms = (float(v) for v in t.cpu())
    value = world * pairs / (ms * 1e-3) / 1e9
    e2e_value = world * pairs / (e2e_ms * 1e-3) / 1e9

    info = ext.device_info()
    peaks, peak_src = measured_peaks()
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    achieved_tflops = FLOP_PER_PAIR * pairs / (search_ms * 1e-3) / 1e12
    ctas, thr, spl, occ = (ctypes.c_int() for _ in range(4))
    lib.rma_chamfer_search_config(B, N, M, ctypes.byref(ctas), ctypes.byref(thr), ctypes.byref(spl), ctypes.byref(occ))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: chamfer fwd+bwd of loss_on_cd, B={B} per GPU, N=M={N}, uniform cube, "
                               f"seed {WORKLOADS[name][3]} (SURVEY.md §8d)",
                   "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "256 MiB flush between timed steps (inputs are 1.5 MiB, L2-resident by nature)",
                   "cuda_graph": bool(use_graph),
                   "search_launch": {"kernel": search_kernel, "ctas": ctas.value, "threads": thr.value,
                                     ("tiles64_per_cta" if filt == 1 else "column_splits"): spl.value,
                                     "ctas_per_sm": occ.value},
                   "search_path": ("filter (expanded form, 3 FFMA + 1 FADD per pair) + exact difference-form resolve"
                                   if filt == 1 else "exact difference form (6 FP32 ops per pair)"),
                   "guard_slow_path": ({"rows": int(stats[0]), "columns": int(stats[1]),
                                        "row_block_rescans": int(stats[2]), "of_items": B * (N + M)}
                                       if filt == 1 else None)},
        "fp32_fma_peak_frac_step": FLOP_PER_PAIR * pairs / (ms * 1e-3) / 1e12 / peak_tflops,
        "roofline": {"bound": "fp32_fma", "kernel": search_kernel, "achieved": achieved_tflops,
                     "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch, from the ncu --set full
                     # capture of the C2 step (profiles/r01p_ncu_full_selected.csv): 2.12 MB read + 0.51 MB written
                     # (compute-bound: inputs and records live in L2); not re-measured for other workloads
                     "traffic": 2.63e6 if (name == "C2" and filt == 1) else None, "kernel_ms": search_ms,
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
        "clocks": clocks,
        "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
        "step_ms_min_max": [min(step_ms), max(step_ms)],
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample_b = min(B, 4)
        sec = cpu_reference_time(x_host, y_host, sample_b, reps=3, threads=threads)
        line["cpu_baseline"] = {
            "value": sample_b * N * M / sec / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample_b} of {B} batch elements, best of 3, torch CPU fp32 (reference formulation: "
                      f"expanded-form [B,N,M] matrix + autograd)"}
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
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
