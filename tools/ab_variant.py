"""A/B harness for compile-time kernel variants.

  build (here, no GPU):   python tools/ab_variant.py build relay -DMMF_TILE2_RELAY=1
                          -> multimodalfusion_b200/libmmf_b200_relay.so (git-ignored, travels with gpurun)
  run (on the GPU box):   python tools/ab_variant.py run default relay [--tests tests/test_gpu_parity.py] [--lanes 1,2]
                          per variant: the GPU parity tests through that library, then bench.py's device-resident step
                          time (MMF_BENCH_QUICK=1) per lane count; prints one table. "default" = the in-tree library.
  clean:                  python tools/ab_variant.py clean

One gpurun call:  gpurun --timeout 600 -- 'python tools/ab_variant.py run default relay > gpurun_out/ab.txt 2>&1; tail -20 gpurun_out/ab.txt'
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodalfusion_b200")
sys.path.insert(0, ROOT)


def lib_path(name):
    return os.path.join(PKG, "libmmf_b200.so" if name == "default" else f"libmmf_b200_{name}.so")


def build(name, defines):
    from multimodalfusion_b200._lib import NVCC_FLAGS
    if name == "default":
        raise SystemExit("the default library is built by __graft_entry__.build()")
    cmd = ["nvcc", *NVCC_FLAGS, *defines, "-o", lib_path(name), os.path.join(PKG, "csrc", "capi.cu")]
    print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    print("built", lib_path(name), os.path.getsize(lib_path(name)), "bytes")


def run(names, tests, lanes):
    rows = []
    for name in names:
        if not os.path.exists(lib_path(name)):
            raise SystemExit(f"{lib_path(name)} missing: build it first")
        env = dict(os.environ, MMF_LIB_PATH=lib_path(name))
        t = subprocess.run([sys.executable, "-m", "pytest", *tests, "-x", "-q", "-m", "gpu"], cwd=ROOT, env=env,
                           capture_output=True, text=True)
        verdict = (t.stdout.strip().splitlines() or ["?"])[-1]
        row = {"variant": name, "tests": verdict}
        if t.returncode != 0:
            print(t.stdout[-3000:], t.stderr[-2000:], sep="\n")
        for n in lanes:
            b = subprocess.run([sys.executable, "bench.py", "--steps", "64", "--warmup", "8"], cwd=ROOT,
                               env=dict(env, MMF_BENCH_QUICK="1"), capture_output=True, text=True)
            try:
                line = json.loads([l for l in b.stdout.splitlines() if l.startswith("{")][-1])
                row[f"us/step @{n} lane(s)"] = round(line["ms_per_step"] * 1e3, 2)
            except (IndexError, ValueError):
                row[f"us/step @{n} lane(s)"] = "failed"
                print(b.stdout[-1500:], b.stderr[-1500:], sep="\n")
        rows.append(row)
        print(json.dumps(row), flush=True)
    keys = list(rows[0])
    print("\n| " + " | ".join(keys) + " |\n|" + "---|" * len(keys))
    for r in rows:
        print("| " + " | ".join(str(r.get(k, "")) for k in keys) + " |")


def main():
    if len(sys.argv) < 2 or sys.argv[1] not in ("build", "run", "clean"):
        raise SystemExit(__doc__)
    if sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == "clean":
        for f in glob.glob(os.path.join(PKG, "libmmf_b200_*.so")):
            os.remove(f)
            print("removed", f)
    else:
        args, tests, lanes = sys.argv[2:], ["tests/test_gpu_parity.py"], [1, 2]
        names = []
        while args:
            a = args.pop(0)
            if a == "--tests":
                tests = args.pop(0).split(",")
            elif a == "--lanes":
                lanes = [int(v) for v in args.pop(0).split(",")]
            else:
                names.append(a)
        run(names or ["default"], tests, lanes)


if __name__ == "__main__":
    main()
