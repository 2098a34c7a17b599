"""Small fixed workload for ncu: N frames x 48 units through vi_inspect_batch, twice."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vi_b200
from vi_b200 import synth
from vi_b200.grid import Grid, generate_grid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
boxes = generate_grid((251, 232, 316, 315), 4, 6, 2, 1, 133, 136, 252, 0)
uniq = [synth.make_frame(s, [b for b, _ in boxes]) for s in range(min(n, 4))]
frames = np.stack([uniq[i % len(uniq)] for i in range(n)])
insp = vi_b200.Inspector(0)
insp.configure(Grid(boxes=boxes), is_reference=True)
d = torch.from_numpy(frames).cuda()
for _ in range(2):
    rec, _, _ = insp.inspect_batch(d)
    torch.cuda.synchronize()
rec = rec.cpu().numpy().view(vi_b200.RECORD_DTYPE).reshape(-1)
print("ng", int((rec['status'] == 1).sum()), "of", len(rec))
