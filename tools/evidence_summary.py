"""Turns one tools/gpu_evidence.sh run (gpurun_out/<R>_*) into profiles/<R>_ncu_step_summary.md (+ raw csv copies).
usage: python tools/evidence_summary.py r02b "<commit / remark>" """
import csv, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, remark = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True, cwd=ROOT).stdout
for f in (f"{R}_launches.csv", f"{R}_warm_traffic.csv"):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))
launch = run("tools/launch_list_summary.py", f"gpurun_out/{R}_launches.csv").split("| `void at::")[0]
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_warm_traffic.csv"))) if len(r) > 10]
h, agg = rows[0], {}
for r in rows[1:]:
    d = dict(zip(h, r))
    k = d["Kernel Name"].split("(")[0].replace("void ", "").replace("mmf::", "")
    agg.setdefault(k, {}).setdefault(d["Metric Name"], []).append(float(d["Metric Value"].replace(",", "")))
out = [f"# {R} — ncu evidence of the fused 3-launch training step ({remark}; 1 x B200, `tools/gpu_evidence.sh {R}`)\n",
       "Each command ran after the same program had exited 0 without ncu (commands: `tools/gpu_evidence.sh`).\n",
       f"## 1. Launch list of `python bench.py --steps 16 --warmup 3` (`profiles/{R}_launches.csv`; serialised cold-cache times: the SHARES are comparable with `stage_us`, the absolute times are not)\n",
       launch,
       f"\n## 2. Warm-cache DRAM / L2 traffic per launch (`--cache-control none`, warm steps of `tools/prof_step.py`; `profiles/{R}_warm_traffic.csv`)\n",
       "| kernel | DRAM read MB | DRAM write MB | L2 bytes MB | ncu us |\n|---|---|---|---|---|"]
tr = tw = 0.0
for k, m in agg.items():
    f = lambda n: sum(m[n]) / len(m[n])
    out.append(f"| `{k}` | {f('dram__bytes_read.sum') / 1e6:.1f} | {f('dram__bytes_write.sum') / 1e6:.1f} | {f('lts__t_bytes.sum') / 1e6:.1f} | {f('gpu__time_duration.sum') / 1e3:.1f} |")
    tr += f("dram__bytes_read.sum"); tw += f("dram__bytes_write.sum")
out.append(f"| **step** | **{tr / 1e6:.1f}** | **{tw / 1e6:.1f}** | | |\n")
out.append(f"Algorithmic bytes of the step: 67.1 MB (the bag x is read twice; weights, gradients and the 56 MB of stash / dG / dU are meant to live in L2). "
           f"Measured DRAM traffic {(tr + tw) / 1e6:.0f} MB = {(tr + tw) / 67.1e6:.2f}x.\n")
out.append("\n## 3. Full capture (`--set full --import-source on`, cold-cache serialised replays: ratios, not times)\n\n```\n" +
           "== ".join([""] + run("tools/ncu_summary.py", f"gpurun_out/{R}_prof_step.ncu-rep").split("== ")[1:4]) + "```\n")
hot = "".join(f"#### {k}\n" + run("tools/ncu_src_hot.py", f"gpurun_out/{R}_prof_step.ncu-rep", k, "0", "12") for k in ("amil_tile2", "amil_hidden", "gemm2_tc"))
out.append("\n## 4. Stall samples per source line (`tools/ncu_src_hot.py`; the single-thread role warps and the warps parked on the final barriers dominate the sample counts)\n\n```\n" + hot + "```\n")
open(os.path.join(P, f"{R}_ncu_step_summary.md"), "w").write("\n".join(out))
print("\n".join(out[:12])[:3000])
