"""Event-timed prep / search / resolve (finalize) stages of one chamfer forward through the C ABI, for sizes where
tools/graph_marginals.py would take minutes.   python tools/time_stages.py C5s4 [reps]   (RMA_B200_LIB selects a build)"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200 import _lib  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C5s4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, N, M, _, _ = WORKLOADS[wl]
dev = torch.device("cuda:0")
x, y = make_workload(wl)
xd, yd = x.to(dev), y.to(dev)
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
opts = _lib.ChamferOpts(_lib.PATH_FILTER, 0, 1 << 40)          # force the filter path whatever its record size
o = ctypes.byref(opts)
wsb = lib.rma_chamfer_workspace_bytes(B, N, M, o)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
xs, ys = xd.stride(), yd.stride()
res = {"workload": wl, "workspace_MiB": wsb / 2**20}
for r in range(reps + 1):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    assert lib.rma_chamfer_prep(P(xd), *xs, P(yd), *ys, B, N, M, P(ws), wsb, o, st) == 0
    ev[1].record()
    assert lib.rma_chamfer_search(B, N, M, P(ws), wsb, o, st) == 0
    ev[2].record()
    assert lib.rma_chamfer_finalize(B, N, M, P(d1), P(d2), P(i1), P(i2), P(ws), wsb, o, st) == 0
    ev[3].record()
    torch.cuda.synchronize()
    if r:
        for k, name in enumerate(("prep_ms", "search_ms", "resolve_ms")):
            res.setdefault(name, []).append(round(ev[k].elapsed_time(ev[k + 1]), 3))
stats = (ctypes.c_uint * 4)()
lib.rma_chamfer_guard_stats(B, N, M, P(ws), o, ctypes.cast(stats, ctypes.c_void_p), st)
res["guard_slow_path"] = {"rows": stats[0], "columns": stats[1], "row_block_rescans": stats[2]}
print(json.dumps(res))
