"""kNN part of the reference's model/network.py, same names and argument meaning, CUDA-only.

    knn(x, k)                 network.py:52-58    x [B,3,N] -> idx [B,N,k] (int64), nearest first, self included
    get_graph_feature(x, k)   network.py:61-83    DGCNN edge features [B,6,N,k] on top of knn (pure indexing, torch)

The reference builds the [B,N,N] matrix `-|xi|^2 + 2 xi.xj - |xj|^2` and takes torch.topk of it; here the matrix is
never materialised (rma_knn, csrc/knn.cu).  Order: (exact fp32 difference-form distance, then index) — torch.topk's
order among equal values is unspecified, and the reference's expanded form can swap near-equal neighbours.
Feature dimensions other than 3 are not on the reference's path (knn is only called on the input cloud,
network.py:264) and raise.
"""
import torch

from .. import ext


def knn(x, k):
    if x.dim() != 3 or x.shape[1] != 3:
        raise RuntimeError(f"knn expects x of shape [B, 3, N], got {tuple(x.shape)}")
    pts = x.detach().transpose(2, 1)                  # [B,N,3] view with strides (3N, 1, N): no copy
    return ext.knn_indices(pts, pts, k).long()        # (batch_size, num_points, k)


def get_graph_feature(x, k=5):
    """x [B,C,N] -> [B,2C,N,k] DGCNN edge features: channels 0..C-1 hold the neighbour's coordinates, C..2C-1 the
    point's own (the value and memory layout of network.py:61-83: a permuted view of a contiguous [B,N,k,2C] tensor)."""
    neighbours = knn(x, k=k)                                             # [B,N,k]
    batch = torch.arange(x.shape[0], device=x.device)[:, None, None]
    points = x.transpose(2, 1)                                           # [B,N,C] view
    theirs = points[batch, neighbours]                                   # [B,N,k,C]
    own = points.unsqueeze(2).expand(-1, -1, neighbours.shape[2], -1)
    return torch.cat((theirs, own), dim=3).permute(0, 3, 1, 2)
