"""Summarise one `ncu --set full` capture of vi_unit_kernel (run here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep UNITS_IN_CAPTURE OUT_PREFIX [--traffic]

Writes OUT_PREFIX.txt (key counters + per-source-line instruction shares) and, with --traffic,
profiles/traffic.json (DRAM bytes per unit and per bench launch of 3,072 units) for bench.py."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__block_size", "launch__grid_size", "sm__cycles_active.avg",
]


def main():
    rep, units, prefix = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr, unit_row, val = rows[0], rows[1], rows[2]
    m = {k: (val[i], unit_row[i]) for i, k in enumerate(hdr)}
    lines = [f"ncu --set full --clock-control none, one launch of vi_unit_kernel over {units} units ({rep})", ""]
    for k in KEYS:
        if k in m:
            lines.append(f"{k:90s} {m[k][0]:>16s} {m[k][1]}")
    lines.append("")
    lines.append("warp stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
    st = [(float(v[0]), k.split("stalled_")[1].split("_per_issue")[0]) for k, v in m.items()
          if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
    for v, k in sorted(st, reverse=True):
        if v >= 0.05:
            lines.append(f"    {k:28s} {v:6.2f}")

    def to_bytes(v, u):
        f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v) * f
    rd = to_bytes(*m["dram__bytes_read.sum"]); wr = to_bytes(*m["dram__bytes_write.sum"])
    per_unit = (rd + wr) / units
    lines += ["", f"DRAM traffic: read {rd/1e6:.1f} MB + write {wr/1e6:.1f} MB = {per_unit:.0f} B per unit "
                  f"(algorithmic: 298,620 B per 316x315 unit = 3 B/px); instructions per unit: "
                  f"{float(m['smsp__inst_executed.sum'][0]) / units:.0f} warp-instructions", ""]
    src = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "0.5"], capture_output=True, text=True).stdout
    lines.append("per-source-line share of executed warp instructions / stall samples (>= 0.5 %):")
    lines.append(src)
    open(prefix + ".txt", "w").write("\n".join(lines))
    print("\n".join(lines[:40]))
    if "--traffic" in sys.argv:
        json.dump({"dram_bytes_per_unit": per_unit, "dram_bytes_per_launch": per_unit * 3072,
                   "launch_units": 3072, "capture_units": units, "source": os.path.basename(prefix) + ".txt"},
                  open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
