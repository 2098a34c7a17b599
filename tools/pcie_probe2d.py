"""Does a strided (crop-only) upload beat uploading whole covered rows?  2-D DMA of 316-byte rows at pitch 4096
against the contiguous copy (development probe; cuda-python's runtime binding for cudaMemcpy2DAsync)."""
import torch
from cuda import cudart
W, rows = 4096, 64 * 1890                      # covered rows of 64 frames
h = torch.empty(rows * W, dtype=torch.uint8, pin_memory=True)
d = torch.empty(rows * W, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def full():
    cudart.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), rows * W, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st.cuda_stream)
def strided(width, ncols, x0s):
    def f():
        for x0 in x0s:
            cudart.cudaMemcpy2DAsync(d.data_ptr() + x0, W, h.data_ptr() + x0, W, width, rows, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st.cuda_stream)
    return f
ms = timed(full); print(f"contiguous {rows*W/1e6:.0f} MB: {ms:.2f} ms ({rows*W/ms/1e6:.1f} GB/s)")
xs = [251 + k * 449 for k in range(4)] + [2166 + k * 449 for k in range(4)]
for width, label in ((316, "316-byte rows (unit crops)"), (336, "336-byte rows (16-byte aligned spans)"), (3578, "one 3578-byte span per row")):
    cols = xs if width < 1000 else [251]
    if width == 336: cols = [x & ~15 for x in xs]
    ms = timed(strided(width, len(cols), cols))
    mb = width * rows * len(cols) / 1e6
    print(f"{label}: {mb:.0f} MB in {ms:.2f} ms ({mb/ms:.1f} GB/s payload)")
