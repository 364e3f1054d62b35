"""Loss.forward + backward (reference model/loss.py:249-355) on the B200 kernels: stage-fused vs the literal per-stage
loop, and the ARAP term alone against the torch composition the reference runs (indexing_neighbor gathers + ~12
elementwise kernels per stage, loss.py:233-243).  Prints one JSON line per measurement.

    python tools/bench_loss_forward.py [B N T]"""
import json
import os
import random
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200.model import loss as L  # noqa: E402


def timed(fn, reps=20):
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


def torch_arap(idx, d, s):
    """the reference's composition (loss.py:113-120, 233-243) in torch, for timing only"""
    bs = idx.size(0)
    id0 = torch.arange(bs, device=d.device).view(-1, 1, 1)
    dn = d[id0, idx] - d.unsqueeze(2)
    sn = s[id0, idx] - s.unsqueeze(2)
    dl = torch.sqrt(torch.mul(dn, dn).sum(dim=-1) + 0.00001)
    sl = torch.sqrt(torch.mul(sn, sn).sum(dim=-1) + 0.00001)
    diff = dl - sl
    return torch.sum(torch.mul(diff, diff)) / bs


def main():
    B, N, T = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 8192, 7)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    source = ((torch.rand(B, N, 3, generator=g) - 0.5)).to(dev)
    target = (source.cpu() + 0.02 * torch.randn(B, N, 3, generator=g)).to(dev)
    deform = [(source.cpu() + 0.01 * (i + 1) * torch.randn(B, N, 3, generator=g)).transpose(1, 2).contiguous().to(dev)
              .requires_grad_(True) for i in range(T)]
    weights = [(torch.rand(B, 1, N, generator=g)).to(dev).requires_grad_(True) for _ in range(T)]
    rigid = [torch.randn(B, 4, 4, generator=g).to(dev).requires_grad_(True) for _ in range(T)]
    phis = [torch.zeros(B, 6, device=dev) for _ in range(T)]
    base = dict(weight_cd=1.0, weight_pd=-1.0, weight_sparse=10.0, weight_depth=-1.0, weight_tran=0.1, weight_mask=-1.0,
                weight_arap=0.01, gamma=1.0, neighbour_num=4, view_divide=5)

    def run(obj):
        def step():
            for t in deform + weights + rigid:
                t.grad = None
            loss, _ = obj(phis, weights, deform, deform, rigid, source, target)
            loss.backward()
        return step

    for name, extra in (("cd+sparse+tran+arap", {}), ("train_sample.py weights (adds depth, mask; 25 views)",
                                                      dict(weight_depth=1.0, weight_mask=0.1))):
        args = SimpleNamespace(**{**base, **extra})
        ms_f = timed(run(L.Loss(args, device=dev, fused=True, rng=random.Random(1))), 10)
        ms_l = timed(run(L.Loss(args, device=dev, fused=False, rng=random.Random(1))), 10)
        print(json.dumps({"what": "Loss.forward + backward", "terms": name, "B": B, "N": N, "T": T,
                          "ms_stage_fused": ms_f, "ms_per_stage_loop": ms_l}), flush=True)

    # ARAP alone: T stages, fused launch vs per-stage launches vs the reference's torch composition
    idx = L.get_neighbor_index(source, 4)
    idx32 = idx.int()

    def arap_fused():
        for t in deform:
            t.grad = None
        L.loss_arap_stages_fused(idx32, deform, source, 1.0).backward()

    def arap_loop():
        for t in deform:
            t.grad = None
        L.loss_arap_stages(idx32, deform, source, 1.0).backward()

    def arap_torch():
        for t in deform:
            t.grad = None
        out = 0
        for i in range(T):
            out = out + L.arap_stage_factor(i) * torch_arap(idx, deform[i].transpose(1, 2), source)
        out.backward()

    a, b, c = timed(arap_fused), timed(arap_loop), timed(arap_torch)
    # kernel-only time and achieved algorithmic bandwidth of the fused launch (n = 4):
    # per (stage, vertex): 12 + 48 B deformation, 12 + 48 B source, 16 B indices, 60 B gradient updates = 196 B
    x_all = L.stack_stages(deform).detach().reshape(T * B, N, 3)
    w = torch.ones(T * B, device=dev)
    from rma_net_b200 import ext

    def kernel_only():
        ext.arap_forward(x_all, source, idx32, w, 1.0 / B, True)

    k = timed(kernel_only, 50)
    print(json.dumps({"what": "ARAP term, T stages fwd+bwd", "B": B, "N": N, "T": T, "ms_one_fused_launch": a,
                      "ms_per_stage_launches": b, "ms_torch_composition_of_the_reference": c,
                      "speedup_vs_torch": c / a, "ms_kernel_plus_two_memsets": k,
                      "algorithmic_GBps_kernel": T * B * N * 196 / (k * 1e-3) / 1e9}), flush=True)
    # neighbour graph: rma_knn vs the reference's [B,V,V] matrix + topk (loss.py:99-110)
    def nb_ours():
        L.get_neighbor_index(source, 4)

    def nb_torch():
        inner = torch.bmm(source, source.transpose(1, 2))
        q = torch.sum(source ** 2, dim=2)
        dist = inner * (-2) + q.unsqueeze(1) + q.unsqueeze(2)
        return torch.topk(dist, k=5, dim=-1, largest=False)[1][:, :, 1:]

    print(json.dumps({"what": "get_neighbor_index (n=4)", "B": B, "N": N, "ms_rma_knn": timed(nb_ours, 10),
                      "ms_torch_matrix_topk": timed(nb_torch, 5)}), flush=True)


if __name__ == "__main__":
    main()
