"""Shared-memory wavefronts (total / excessive) per CUDA source line from an .ncu-rep."""
import csv, io, subprocess, sys
path, kern = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname = func = hdr = None
agg = {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or kern not in (func or "") or len(r) < len(hdr) - 5 or r[2] != "-": continue
    iw, ix, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("L1 Wavefronts Shared Ideal")
    ig = hdr.index("L2 Theoretical Sectors Global")
    try: w, x, idl, g = int(r[iw] or 0), int(r[ix] or 0), int(r[ii] or 0), int(r[ig] or 0)
    except ValueError: continue
    a = agg.setdefault((fname, int(r[0])), [0, 0, 0, 0, r[1]])
    a[0] += w; a[1] += x; a[2] += idl; a[3] += g
tw = sum(a[0] for a in agg.values()) or 1
print("total shared wavefronts", tw, "excessive", sum(a[1] for a in agg.values()))
for (f, ln), a in sorted(agg.items()):
    if a[0] * 200 > tw or a[3] > 0:
        print(f"{f:12s}:{ln:4d} wf {100.0*a[0]/tw:5.1f}%  excess {100.0*a[1]/tw:5.1f}%  glob-sectors {a[3]:>10d}  {a[4].strip()[:80]}")
