"""Where the headline leg's extra ~5 us per step come from: R CUDA graphs replayed in rotation, (a) all on the SAME
input pair, (b) each on its own input pair, against one graph replayed back to back; and one graph holding n steps on
n different pairs.    python tools/rotate_probe.py [R]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200.model.loss import loss_on_cd  # noqa: E402

dev = torch.device("cuda:0")
B, N = 16, 4096
R = int(sys.argv[1]) if len(sys.argv) > 1 else 93
gen = torch.Generator().manual_seed(0)
pairs = [((torch.rand(B, N, 3, generator=gen) * 2 - 1).to(dev).requires_grad_(True),
          (torch.rand(B, N, 3, generator=gen) * 2 - 1).to(dev).requires_grad_(True)) for _ in range(R)]


def make_step(x, y):
    def step():
        x.grad = None; y.grad = None
        loss_on_cd(x, y).backward()
    return step


def capture(fns):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    return g


def timed(graphs, steps_per_graph, K=186):
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(40e6))
        e0.record()
        for k in range(K):
            graphs[k % len(graphs)].replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / K / steps_per_graph)
    return round(best, 2)


out = {}
one = capture([make_step(*pairs[0])])
out["one graph, one pair"] = timed([one], 1)
same = [capture([make_step(*pairs[0])]) for _ in range(R)]
out[f"{R} graphs, all on one pair"] = timed(same, 1)
del same
own = [capture([make_step(*pairs[r])]) for r in range(R)]
out[f"{R} graphs, own pair each"] = timed(own, 1)
del own
n = 31
multi = [capture([make_step(*pairs[r]) for r in range(i, i + n)]) for i in range(0, R - n + 1, n)]
out[f"{len(multi)} graphs of {n} steps on {n} pairs each"] = timed(multi, n, K=12)
print(out, flush=True)
