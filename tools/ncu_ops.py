"""Executed-instruction histogram by opcode (and by source line for a given opcode prefix) from an ncu report.
   python tools/ncu_ops.py rep.ncu-rep [OPCODE_PREFIX]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vision-inspection-system-segmentation-using-classical-computer-vision-_b200", "libvi_b200.so")
rep = sys.argv[1]
pref = sys.argv[2] if len(sys.argv) > 2 else None
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
ia = hdr.index("Instructions Executed")
base = int(data[0][0], 16)
ops, lines = collections.Counter(), collections.Counter()
tot = 0
for r in data:
    off = int(r[0], 16) - base
    fl, txt = seq.get(off, (None, "?"))
    t = txt.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    op = op.split(".")[0]
    n = int(r[ia]); tot += n
    ops[op] += n
    if pref and op.startswith(pref):
        lines[fl] += n
print("total", tot)
for op, n in ops.most_common(40):
    print(f"  {op:12s} {100*n/tot:6.2f}%")
if pref:
    print("lines for", pref)
    for fl, n in lines.most_common(40):
        print(f"  {str(fl):36s} {100*n/tot:6.2f}%")
