"""Print the SASS of vi_unit_kernel with per-instruction executed counts from an ncu report, for source lines
FILE:LO-HI (development helper).   python tools/ncu_sass.py rep.ncu-rep vi_rank.cuh 165 185 [min_exec]"""
import csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vision-inspection-system-segmentation-using-classical-computer-vision-_b200", "libvi_b200.so")
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
min_exec = int(sys.argv[5]) if len(sys.argv) > 5 else 1
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
seq, cur, inker = {}, None, False
for ln in dis.split("\n"):
    if ln.startswith("//---") and ".text." in ln:
        inker = "vi_unit_kernelILb0ELb1ELb0E" in ln
        continue
    if not inker:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(data[0][0], 16)
tot = 0
for r in data:
    off = int(r[0], 16) - base
    fl, txt = seq.get(off, (None, "?"))
    if fl and fl[0] == fname and lo <= fl[1] <= hi and int(r[ia]) >= min_exec:
        print(f"{off:6x} L{fl[1]:4d} {int(r[ia]):9d} {int(r[isamp]):5d}  {txt}")
        tot += int(r[ia])
print("total executed", tot)
