"""Small fixed workload for ncu: the two frame-ingest kernels on 16 frames of 4096x3000, three launches each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vi_b200
insp = vi_b200.Inspector(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
a = torch.empty((n, 3000, 4096, 4), dtype=torch.uint8, device="cuda"); a.random_(0, 256)
u = torch.empty((n, 3000, 4096), dtype=torch.int16, device="cuda"); u.view(torch.uint8).random_(0, 256)
out = torch.empty((n, 3000, 4096), dtype=torch.uint8, device="cuda")
for _ in range(3):
    insp.ingest_argb32(a, out=out)
    insp.ingest_gray16(u, out=out)
torch.cuda.synchronize()
print("ok", int(out.sum()))
