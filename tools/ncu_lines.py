"""Join an ncu report's per-SASS-instruction counters with CUDA source lines.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_pct]

Needs: the .so built with -lineinfo (same build as profiled), cuobjdump, nvdisasm, ncu (all run here, no GPU).
Prints, per source file, every line whose share of executed warp-instructions is >= min_pct,
with its share of stall samples, then per-function totals."""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vision-inspection-system-segmentation-using-classical-computer-vision-_b200", "libvi_b200.so")


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
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
            seq[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][0], 16)
    inst, samp = collections.Counter(), collections.Counter()
    for r in data:
        fl = seq.get(int(r[0], 16) - base)
        inst[fl] += int(r[ia])
        samp[fl] += int(r[isamp])
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"warp instructions executed {ti}, stall samples {ts}")
    srcdir = os.path.join(ROOT, "vision-inspection-system-segmentation-using-classical-computer-vision-_b200", "csrc")
    files = collections.defaultdict(dict)
    for (k, v) in inst.items():
        if k:
            files[k[0]][k[1]] = v
    for f in sorted(files, key=lambda f: -sum(files[f].values())):
        tot = sum(files[f].values())
        print(f"\n=== {f}: {100 * tot / ti:.1f}% of instructions")
        path = os.path.join(srcdir, f)
        src = open(path).read().split("\n") if os.path.exists(path) else []
        for ln in sorted(files[f]):
            p = 100 * files[f][ln] / ti
            if p >= min_pct:
                s = 100 * samp[(f, ln)] / ts
                text = src[ln - 1].strip()[:100] if ln - 1 < len(src) else ""
                print(f"  {ln:4d} {p:5.2f}% inst {s:5.2f}% stall | {text}")


if __name__ == "__main__":
    main()
