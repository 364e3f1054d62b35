"""tests/golden/loss_forward.npz: outputs of the REFERENCE's own Loss.forward and term methods on CPU (this container).

model/loss.py cannot be imported here (trimesh, the compiled point_render extension, cuda:0 buffers in __init__), but
the pieces this fixture needs are self-contained torch and are exec'd VERBATIM from /root/reference/model/loss.py:
  45-70    compute_sqrdis_map, chamfer_dist          99-120  get_neighbor_index, indexing_neighbor
  200-243  loss_on_cd / pd / sparse / tran / arap    249-355 Loss.forward (the stage loops, gamma and the arap factors)
as the body of a stand-in class whose __init__ only sets `args` and `trans_mask` (loss.py:134-135) on the CPU.
weight_depth = weight_mask = -1 (those terms need the CUDA extension; they are pinned by tests/golden/exfunc_*.npz).
   python oracle/gen_golden_loss.py"""
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/model/loss.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _lines(first, last):
    with open(REF) as f:
        return "".join(f.readlines()[first - 1:last])


def reference_loss_class():
    ns = {"torch": torch, "nn": nn}
    exec(compile(_lines(45, 70) + _lines(99, 120), REF, "exec"), ns)
    body = ("class RefLoss:\n"
            "    def __init__(self, args):\n"
            "        self.args = args\n"
            "        self.trans_mask = torch.zeros([1,4,4],dtype=torch.float32,requires_grad=False)\n"
            "        self.trans_mask[:, 0:3, 3]+=1.0\n"
            "    def __call__(self, *a):\n"
            "        return self.forward(*a)\n" + _lines(200, 243) + _lines(249, 355))
    exec(compile(body, REF, "exec"), ns)
    return ns["RefLoss"], ns


def main():
    RefLoss, ns = reference_loss_class()
    args = SimpleNamespace(weight_cd=1.0, weight_pd=0.5, weight_sparse=10.0, weight_depth=-1.0, weight_tran=0.1,
                           weight_mask=-1.0, weight_arap=0.01, gamma=0.8, neighbour_num=4, view_divide=2)
    B, N, T = 2, 160, 7                                   # T = 7 walks all three arap factor tiers (loss.py:337-343)
    g = torch.Generator().manual_seed(11)
    source = torch.rand(B, N, 3, generator=g) * 2 - 1
    target = source + 0.1 * torch.randn(B, N, 3, generator=g)
    deform = [(source + 0.05 * (i + 1) * torch.randn(B, N, 3, generator=g)).transpose(1, 2).contiguous()
              .requires_grad_(True) for i in range(T)]                         # [B,3,N] as the network returns them
    weights = [(torch.rand(B, 1, N, generator=g) - 0.3).requires_grad_(True) for _ in range(T)]
    rigid = [torch.randn(B, 4, 4, generator=g).requires_grad_(True) for _ in range(T)]
    phis = [torch.randn(B, 6, generator=g) for _ in range(T)]
    loss_obj = RefLoss(args)
    loss, stages = loss_obj(phis, weights, deform, deform, rigid, source, target)
    loss.backward()
    nb = ns["get_neighbor_index"](source, args.neighbour_num)
    single = {
        "pd": loss_obj.loss_on_pd(deform[0].transpose(1, 2), target),
        "sparse": loss_obj.loss_on_sparse(weights[0]),
        "tran": loss_obj.loss_on_tran(rigid[0]),
        "arap": loss_obj.loss_on_arap(nb, deform[0].transpose(1, 2), source),
        "cd": loss_obj.loss_on_cd(deform[0].transpose(1, 2), target),
    }
    out = dict(source=source.numpy(), target=target.numpy(),
               deform=np.stack([d.detach().numpy() for d in deform]),
               weights=np.stack([w.detach().numpy() for w in weights]),
               rigid=np.stack([r.detach().numpy() for r in rigid]),
               phis=np.stack([p.numpy() for p in phis]),
               args=np.array([args.weight_cd, args.weight_pd, args.weight_sparse, args.weight_depth, args.weight_tran,
                              args.weight_mask, args.weight_arap, args.gamma, args.neighbour_num], np.float64),
               neighbour_idx=nb.numpy().astype(np.int32),
               ref_loss=np.float32(loss.item()), ref_stages=np.array([float(s) for s in stages], np.float32),
               ref_grad_deform=np.stack([d.grad.numpy() for d in deform]),
               ref_grad_weights=np.stack([w.grad.numpy() for w in weights]),
               ref_grad_rigid=np.stack([r.grad.numpy() for r in rigid]))
    for k, v in single.items():
        out["ref_single_" + k] = np.float32(v.item())
    # euler2rot (loss.py:73-95, exec'd verbatim) on seeded angles, incl. the view grid's range [-pi, pi)
    ens = {"torch": torch}
    exec(compile(_lines(73, 95), REF, "exec"), ens)
    angles = (torch.rand(40, 3, generator=g) * 2 - 1) * float(np.pi)
    out["euler_angles"] = angles.numpy()
    out["ref_euler2rot"] = ens["euler2rot"](angles).numpy()
    np.savez_compressed(os.path.join(GOLD, "loss_forward.npz"), **out)
    print("loss", loss.item(), "stages", [float(s) for s in stages])


if __name__ == "__main__":
    main()
