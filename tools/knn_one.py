"""One kNN search at a given size (developer timing / ncu target): python tools/knn_one.py [N] [d]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from infercnvpy_b200.pp._neighbors import knn_device

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = int(sys.argv[2]) if len(sys.argv) > 2 else 50
g = torch.Generator(device="cuda"); g.manual_seed(0)
P = torch.randn((N, d), generator=g, device="cuda") + 3.0 * torch.randn((16, d), generator=g, device="cuda")[torch.randint(0, 16, (N,), generator=g, device="cuda")]
for _ in range(2):
    idx, dist = knn_device(P, 15)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); idx, dist = knn_device(P, 15); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
flop_exec = 2.0 * N * N * ((d + 7) // 8 * 8) * 3
print(f"knn N={N} d={d}: {ms:.3f} ms, {N / ms * 1e3:.3e} queries/s, executed {flop_exec / ms / 1e9:.1f} TFLOP/s (3xTF32), algorithmic {2.0 * N * N * d / ms / 1e9:.1f} TFLOP/s")
