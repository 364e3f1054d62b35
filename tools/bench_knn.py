"""kNN: rma_knn vs the reference formulation (network.py:52-58: [B,N,N] expanded-form matrix + torch.topk) on the same
GPU, B=16, k=5.   python tools/bench_knn.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200.model.network import knn  # noqa: E402


def ref_knn(x, k):
    inner = -2 * torch.matmul(x.transpose(2, 1).contiguous(), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pairwise_distance = -xx - inner - xx.transpose(2, 1).contiguous()
    return pairwise_distance.topk(k=k, dim=-1)[1]


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


dev = torch.device("cuda:0")
for N in (2048, 8192):
    g = torch.Generator().manual_seed(N)
    x = (torch.rand(16, 3, N, generator=g) * 2 - 1).to(dev)
    ours = timed(lambda: knn(x, 5))
    ref = timed(lambda: ref_knn(x, 5))
    same = (knn(x, 5) == ref_knn(x, 5)).all(-1).float().mean().item()
    print(json.dumps({"config": f"knn B=16 N={N} k=5", "ms_ours": ours, "ms_reference_formulation_gpu": ref,
                      "G_pairs_per_s_ours": 16 * N * N / ours / 1e6, "rows_identical_to_reference": same}))
