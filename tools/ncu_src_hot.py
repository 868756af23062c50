"""Stall samples of one kernel of an .ncu-rep aggregated per CUDA source line (needs -lineinfo + --import-source on).
usage: ncu_src_hot.py REP KERNEL_REGEX [launch_index] [top_n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern,
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
cur, hdr, agg = None, None, []
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        # duplicate header names ("Source" twice): index by position
        g = lambda name: r[hdr.index(name)]
        try:
            s = int(g("# Samples"))
        except ValueError:
            continue
        agg.append((s, cur.split("/")[-1], r[0], r[1][:120], int(g("Instructions Executed") or 0), r))
tot = sum(a[0] for a in agg) or 1
print("total samples", tot)
stk = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for s, f, l, src, ie, r in sorted(agg, key=lambda a: -a[0])[:top_n]:
    top = sorted(((int(r[i] or 0), k) for i, k in stk), reverse=True)[:3]
    print(f"{s:6d} {100 * s / tot:5.1f}% inst={ie:8d} {f}:{l:>4s} {' '.join(f'{k}={v}' for v, k in top if v)} | {src.strip()}")
