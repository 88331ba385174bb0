"""Per CUDA source line: executed warp-instructions and stall samples from an .ncu-rep (needs -lineinfo).
Usage: python scripts/ncu_by_line.py x.ncu-rep [kernel-substring] [top N]"""
import csv
import io
import subprocess
import sys


def main(path, kern="", top=60):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, func, hdr = None, None, None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or kern not in (func or ""):
            continue
        if len(r) < len(hdr) - 5 or r[2] != "-":      # keep the per-line rows only (address column "-")
            continue
        ie, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            n, s = int(r[ie] or 0), int(r[ismp] or 0)
        except ValueError:
            continue
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += n
        a[1] += s
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    print(f"total inst {tot_i}  samples {tot_s}")
    for (f, ln), a in sorted(agg.items()):
        if a[0] * 1000 > tot_i or a[1] * 1000 > tot_s:
            print(f"{f:14s}:{ln:4d} inst {100.0 * a[0] / tot_i:5.2f}%  stall {100.0 * a[1] / tot_s:5.2f}%  {a[2].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 60)
