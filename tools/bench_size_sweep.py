"""Chamfer search kernel and fused fwd+bwd step versus problem size, filter path and exact path, one GPU.
    python tools/bench_size_sweep.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200 import _lib, ext  # noqa: E402
from rma_net_b200.model.loss import loss_on_cd  # noqa: E402

PEAK_FMA = 148 * 128 * 1.965e9
dev = torch.device("cuda:0")
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


for (B, N) in ((64, 1024), (16, 2048), (16, 4096), (16, 8192), (4, 16384), (1, 65536)):
    g = torch.Generator().manual_seed(N)
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev).requires_grad_(True)
    y = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev).requires_grad_(True)
    pairs = B * N * N
    row = {"B": B, "N=M": N, "pairs": pairs}
    for name, path in (("filter", 2), ("exact", 1)):
      with ext.chamfer_options(path=path):
        o = ctypes.byref(_lib.ChamferOpts(path, 0, 0))
        wsb = lib.rma_chamfer_workspace_bytes(B, N, N, o)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        xs, ys = x.stride(), y.stride()
        lib.rma_chamfer_prep(P(x), *xs, P(y), *ys, B, N, N, P(ws), wsb, o, st)
        t = timed(lambda: lib.rma_chamfer_search(B, N, N, P(ws), wsb, o, st), 20)
        row[f"{name}_search_us"] = round(t * 1e6, 1)
        row[f"{name}_search_frac_fp32_peak"] = round(6 * pairs / t / PEAK_FMA, 3)

        def step():
            x.grad = None
            y.grad = None
            loss_on_cd(x, y).backward()
        sg = torch.cuda.Stream()
        sg.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(sg):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(sg)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        t = timed(graph.replay, 20)
        row[f"{name}_fwd_bwd_us"] = round(t * 1e6, 1)
        row[f"{name}_fwd_bwd_frac_fp32_peak"] = round(6 * pairs / t / PEAK_FMA, 3)
        row[f"{name}_workspace_MiB"] = round(wsb / 2 ** 20, 1)
    print(json.dumps(row), flush=True)
