"""SASS audit: in every kernel that executes griddepcontrol.wait (ACQBULK), no global load may be scheduled before it.
nvcc hoists ld.global.nc / const __restrict__ loads above the wait (they are "immutable" to it) unless the address depends
on a volatile asm placed after the wait (pdl_fresh, csrc/mmf_ptx.cuh). Prints offenders; exit code 1 if any."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def audit(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, seen_wait, pending, bad, kernels = None, False, [], [], 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur, seen_wait, pending = m.group(1), False, []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m or cur is None:
            continue
        op = m.group(2)
        if op.startswith("ACQBULK") and not seen_wait:
            seen_wait = True
            kernels += 1
            bad += [(cur, off, o) for off, o in pending]
        elif not seen_wait and (op.startswith("LDG") or op.startswith("LD.E") or op.startswith("ATOMG") or op.startswith("UTMALDG")):
            pending.append((m.group(1), op))
    return kernels, bad


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodalfusion_b200", "libmmf_b200.so")
    n, bad = audit(lib)
    print(f"{n} kernels execute griddepcontrol.wait; {len(bad)} global loads scheduled before it")
    for k, off, op in bad:
        print("  ", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:100], off, op)
    sys.exit(1 if bad else 0)
