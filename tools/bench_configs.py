"""Secondary configurations of BASELINE.json (not the headline bench line): prints one JSON line per config.

  C3  recurrent alignment step: 8 iterations of se3 exp-map warp + chamfer, B=16, N=M=8192, fwd+bwd
      (network.py:652-655 + loss.py:260-267 semantics, gamma = 0.8), plus achieved HBM GB/s of the warp kernels
  C4  throughput sweep B=512 (split over the ranks), N=M=2048, fwd+bwd
  C5  single huge pair N=M=2^20, source-sharded over the ranks with the NCCL min all-reduce merge, fwd+bwd

  python tools/bench_configs.py C3            (1 GPU)
  torchrun --nproc-per-node 8 ... tools/bench_configs.py C4 C5
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rma_net_b200 import parallel  # noqa: E402
from rma_net_b200.model.exfunc.functions import Warp  # noqa: E402
from rma_net_b200.model.loss import loss_on_cd  # noqa: E402

PEAK_FMA = 148 * 128 * 1.965e9


def timed(fn, dev, reps, world):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfgs = sys.argv[1:] or ["C3"]
    out = []
    if "C3" in cfgs:
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
            loss, prev = 0, None
            for k in range(T):
                out_k, _, _ = warp(phis[k], src, None if k == 0 else ws[k], prev)
                loss = loss + gamma ** (T - k - 1) * loss_on_cd(out_k, tgt)
                prev = out_k
            loss.backward()
            return loss

        ms = timed(step, dev, 10, world)

        # the same step the way the reference structures it (network forward produces all stage outputs, then
        # Loss.forward loops over them, loss.py:260-267): 8 warps, then ONE batched chamfer call over the T stages
        from rma_net_b200.model.loss import loss_cd_stages_fused

        def step_fused():
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

        ms_fused = timed(step_fused, dev, 10, world)

        # the same two steps replayed as CUDA graphs: eager, the 8 warp forward / backward pairs and the autograd glue
        # are host-bound (a warp fwd+bwd pair costs ~0.2 ms of Python against ~10 us of device time)
        def graphed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            return timed(gr.replay, dev, 20, world)

        ms_graph = graphed(step)
        ms_fused_graph = graphed(step_fused)
        pairs = T * B * N * N
        # warp kernels alone (HBM-bound): fwd 40 B/point, bwd reads src+prev+w+go, writes grad_w
        prev = src.clone()

        def warps_only():
            o, _, _ = warp(phis[1], src, ws[1], prev)
            o.backward(src)

        wms = timed(warps_only, dev, 50, world)
        out.append({"config": "C3", "what": "8 x (se3 warp + chamfer) fwd+bwd, B=16 per GPU, N=M=8192", "n_gpus": world,
                    "ms_per_step": ms, "G_pairs_per_s": world * pairs / ms / 1e6,
                    "fp32_fma_peak_frac": 6 * pairs / (ms * 1e-3) / PEAK_FMA,
                    "ms_per_step_stage_fused": ms_fused,
                    "fp32_fma_peak_frac_stage_fused": 6 * pairs / (ms_fused * 1e-3) / PEAK_FMA,
                    "ms_per_step_cuda_graph": ms_graph,
                    "fp32_fma_peak_frac_cuda_graph": 6 * pairs / (ms_graph * 1e-3) / PEAK_FMA,
                    "ms_per_step_stage_fused_cuda_graph": ms_fused_graph,
                    "fp32_fma_peak_frac_stage_fused_cuda_graph": 6 * pairs / (ms_fused_graph * 1e-3) / PEAK_FMA,
                    "warp_fwd_bwd_pair_ms": wms,
                    "warp_fwd_bwd_algorithmic_GBps": B * N * (40 + 44) / (wms * 1e-3) / 1e9})
    if "C4" in cfgs:
        Bg, N = 512, 2048
        s, e = parallel.shard_range(Bg, rank, world)
        g = torch.Generator().manual_seed(3)
        x = (torch.rand(Bg, N, 3, generator=g) * 2 - 1)[s:e].to(dev).requires_grad_(True)
        y = (torch.rand(Bg, N, 3, generator=g) * 2 - 1)[s:e].to(dev).requires_grad_(True)

        def step4():
            x.grad = None
            y.grad = None
            loss_on_cd(x, y).backward()

        ms_eager = timed(step4, dev, 20, world)
        # the shard is small at 8 GPUs (64 batch elements = 85 us of GPU work): replay the step as a CUDA graph, as
        # bench.py does, so that the Python / launch overhead (~250 us eager) does not hide the scaling
        sg = torch.cuda.Stream()
        sg.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(sg):
            for _ in range(3):
                step4()
        torch.cuda.current_stream().wait_stream(sg)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step4()
        ms = timed(graph.replay, dev, 50, world)
        pairs = Bg * N * N
        out.append({"config": "C4", "what": "B=512 batch-sharded, N=M=2048, fwd+bwd, strong scaling (CUDA graph)",
                    "n_gpus": world, "ms_per_step": ms, "ms_per_step_eager": ms_eager, "G_pairs_per_s": pairs / ms / 1e6,
                    "fp32_fma_peak_frac_per_gpu": 6 * pairs / world / (ms * 1e-3) / PEAK_FMA})
    if "C5" in cfgs:
        N = M = 1 << 20
        s, e = parallel.shard_range(N, rank, world)
        g = torch.Generator().manual_seed(4)
        xfull = torch.rand(1, N, 3, generator=g) * 2 - 1
        y = (torch.rand(1, M, 3, generator=g) * 2 - 1).to(dev).requires_grad_(True)
        xl = xfull[:, s:e].to(dev).requires_grad_(True)

        def step5():
            xl.grad = None
            y.grad = None
            d1, d2 = parallel.chamfer_dist_source_sharded(xl, y, s)
            (d1.sum() + d2.sum() / world).backward()

        ms = timed(step5, dev, 3, world)
        pairs = N * M
        out.append({"config": "C5", "what": "single pair N=M=2^20, source-sharded, NCCL min all-reduce merge, fwd+bwd",
                    "n_gpus": world, "ms_per_step": ms, "G_pairs_per_s": pairs / ms / 1e6,
                    "fp32_fma_peak_frac_per_gpu": 6 * pairs / world / (ms * 1e-3) / PEAK_FMA})
    if rank == 0:
        for o in out:
            print(json.dumps(o), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
