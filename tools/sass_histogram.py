"""cuobjdump -sass opcode histogram of the hot kernels in the in-tree library (proof of tcgen05 / TMA / TMEM use).
usage: python tools/sass_histogram.py [substring filters ...] > profiles/rNN_sass_histogram.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("MMF_LIB_PATH") or os.path.join(ROOT, "multimodalfusion_b200", "libmmf_b200.so")
want = sys.argv[1:] or ["amil_tile2_kernel<512, 384, true, 0, true, false>", "amil_tile2_kernel<512, 384, true, 0, false, false>",
                        "amil_hidden_fused_kernel<512, 384, true, false, true>", "gemm2_tc_kernel<1, 1, 1, 512>",
                        "p2p_allreduce_sum_kernel", "amil_tile2_kernel<256, 256, true, 0, true, false>"]
KEYS = ("LDGMC", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU",
        "FFMA2", "FMUL2", "FADD2", "REDG", "ATOMG", "LDG", "STG", "LDS", "STS", "HFMA2", "F2FP", "SHFL", "LD.E", "ST.E")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, hist = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
names = list(hist)
dem = dict(zip(names, subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()))
print(f"# SASS opcode histogram — {os.path.relpath(LIB, ROOT)} (`cuobjdump -sass`, sm_100a)\n")
print("Families: UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add,\n"
      "LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier, FFMA2 / FMUL2 = packed f32x2, LDGMC = multimem.ld_reduce (NVLS).\n")
for k in sorted(names, key=lambda n: dem[n]):
    d = dem[k]
    if not any(w.replace("true", "(bool)1").replace("false", "(bool)0") in d or w in d.replace("(int)", "").replace("(bool)1", "true").replace("(bool)0", "false") for w in want):
        continue
    h = hist[k]
    fam = collections.Counter()
    for op, c in h.items():
        for kk in KEYS:
            if op.startswith(kk):
                fam[op if kk in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "UTCBAR", "MUFU") else kk] += c
                break
    short = d.replace("(int)", "").replace("(bool)1", "true").replace("(bool)0", "false").split("(CUtensorMap")[0].split("(mmf::")[0]
    print(f"## `{short}` — {sum(h.values())} instructions")
    print("* tensor / TMA / TMEM / sync:", ", ".join(f"{a} x{b}" for a, b in sorted(fam.items())))
    print("* top opcodes:", ", ".join(f"{a} x{b}" for a, b in h.most_common(10)), "\n")
