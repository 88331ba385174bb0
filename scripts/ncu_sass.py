"""SASS listing of a kernel from an .ncu-rep with executed count, stall samples and the dominant stall reasons per instruction.
Usage: python scripts/ncu_sass.py x.ncu-rep [min_exec] > listing.txt"""
import csv, io, subprocess, sys
path = sys.argv[1]
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        stall = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
        continue
    ex = int(r[hdr.index("Instructions Executed")] or 0)
    smp = int(r[hdr.index("# Samples")] or 0)
    if ex < min_exec:
        continue
    top = sorted(((int(r[i] or 0), n) for i, n in stall), reverse=True)[:2]
    tops = " ".join(f"{n}:{v}" for v, n in top if v)
    print(f"{r[0][-5:]} ex {ex:8d} smp {smp:5d}  {r[1][:70]:70s} {tops}")
