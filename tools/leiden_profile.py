"""Where the time of leiden_device goes at a given graph size (developer aid): python tools/leiden_profile.py [N]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from infercnvpy_b200.pp._neighbors import neighbors_device
from infercnvpy_b200.tl import _leiden as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
lab = torch.randint(0, 12, (N,), generator=g, device=dev)
Y = 3.0 * torch.randn((12, 50), generator=g, device=dev)[lab] + torch.randn((N, 50), generator=g, device=dev)
gr = neighbors_device(Y, 15)
r, c, w = gr["coo"]
order = torch.argsort(r * N + c)
indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
indptr[1:] = torch.cumsum(torch.bincount(r, minlength=N), 0)
indices, w = c[order].to(torch.int32), w[order]
torch.cuda.synchronize()
# wrap the pieces
acc = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g2(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = f(*a, **k)
        torch.cuda.synchronize(); acc[key] = acc.get(key, 0.0) + time.perf_counter() - t0; acc[key + "_n"] = acc.get(key + "_n", 0) + 1
        return out
    setattr(obj, name, g2)
_orig_sweeps = L._sweeps
def _timed_sweeps(lib, graph, kdeg, two_m, comm, bound, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = _orig_sweeps(lib, graph, kdeg, two_m, comm, bound, *a, **k)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    acc["sweeps"] = acc.get("sweeps", 0.0) + dt
    acc.setdefault("calls", []).append((graph[0].numel() - 1, graph[1].numel(), "refine" if bound is not None else "move", out[2], round(dt * 1e3, 2)))
    return out
L._sweeps = _timed_sweeps
wrap(torch, "unique", "torch.unique")
for tol in (0.0, 1e-5, 1e-4, 1e-3):
    for rep in range(2):
        acc.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        labels = L.leiden_device(indptr, indices, w, move_tol=tol)
        torch.cuda.synchronize(); total = time.perf_counter() - t0
    q = L.modularity_device(indptr, indices, w, labels)
    purity = 0.0
    joint = torch.zeros((int(labels.max()) + 1, 12), dtype=torch.int64, device=dev)
    joint.index_put_((labels, lab), torch.ones_like(lab), accumulate=True)
    purity = float(joint.max(dim=1).values.sum()) / N
    print(f"move_tol={tol:g}: total {total:.3f} s, clusters {int(labels.max()) + 1}, quality {q:.6f}, purity {purity:.4f}, sweeps {acc['sweeps']:.3f} s")
    for c in acc["calls"][:4]:
        print("   nodes %8d edges %9d %-6s sweeps %3d  %8.2f ms" % c)
