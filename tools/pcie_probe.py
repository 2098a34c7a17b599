"""PCIe copy rates of the box (pinned host memory), separately and both directions at once: the floor of the
host-buffer batch call (development helper).   python tools/pcie_probe.py"""
import torch
n_up, n_dn = 495452160, 611770368          # bytes per bench step: covered frame rows up, masks + records down
hu = torch.empty(n_up, dtype=torch.uint8, pin_memory=True); du = torch.empty(n_up, dtype=torch.uint8, device="cuda")
hd = torch.empty(n_dn, dtype=torch.uint8, pin_memory=True); dd = torch.empty(n_dn, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): du.copy_(hu, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2): hd.copy_(dd, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, up, dn in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
    run(up, dn, 2)
    ms = run(up, dn)
    print(f"{name:9s} {ms:7.2f} ms per step-equivalent  ({(n_up*up + n_dn*dn)/ms/1e6:.1f} GB/s aggregate)")
