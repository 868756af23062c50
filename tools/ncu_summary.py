"""Print a compact per-kernel table from an .ncu-rep (raw page CSV)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
stall = [i for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:70], d.get("Grid Size"), d.get("Block Size"))
    for k in keys:
        for h in hdr:
            if h == k or (k.endswith("_tensor") and h.startswith(k)):
                print(f"   {h:75s} {d[h]:>16s} {units[hdr.index(h)]}")
    st = sorted(((float(r[i] or 0), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i in stall), reverse=True)[:6]
    print("   stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st))
