"""tests/golden/knn_*.npz: outputs of the REFERENCE's own kNN code on CPU (this container):
`knn` is exec'd verbatim from /root/reference/model/network.py:52-58 and `get_neighbor_index` from
/root/reference/model/loss.py:99-110 (both are self-contained torch).  Inputs: the samples/ OBJ vertex sets already
stored in tests/golden/chamfer_samples*.npz and a seeded uniform cloud.   python oracle/gen_golden_knn.py"""
import os

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _exec_lines(path, first, last, ns):
    with open(path) as f:
        src = "".join(f.readlines()[first - 1:last])
    exec(compile(src, path, "exec"), ns)


def main():
    ns = {"torch": torch}
    _exec_lines(os.path.join(REF, "model", "network.py"), 52, 58, ns)
    _exec_lines(os.path.join(REF, "model", "loss.py"), 99, 110, ns)
    knn, get_neighbor_index = ns["knn"], ns["get_neighbor_index"]
    clouds = {}
    for name in ("samples1", "samples3"):
        g = np.load(os.path.join(GOLD, f"chamfer_{name}.npz"))
        clouds[name] = np.stack([g["x"][0], g["y"][0]])                       # [2, N, 3]: source and target
    gen = torch.Generator().manual_seed(9)
    clouds["uniform"] = (torch.rand(2, 1500, 3, generator=gen) * 2 - 1).numpy()
    for name, pts in clouds.items():
        t = torch.from_numpy(pts)
        idx5 = knn(t.transpose(2, 1).contiguous(), 5).numpy()                 # network.py:63 default k = 5
        nb4 = get_neighbor_index(t, 4).numpy()                                # config.py:128 default neighbour_num = 4
        np.savez_compressed(os.path.join(GOLD, f"knn_{name}.npz"), pts=pts, ref_knn5=idx5.astype(np.int32),
                            ref_neighbor4=nb4.astype(np.int32))
        print(name, pts.shape, "ok")


if __name__ == "__main__":
    main()
