"""Ablation builds of logmel_kernel (results in profiles/r02_logmel_ablation.txt).
Patches a COPY of csrc/logmel.cu (exact-text replacements, asserted) with switches selected by -DMODFX_EXP=<bits>:
  1 no mel taps (one tap per band)   2 no FFT passes   4 no global stores   16 stores as 128-byte segments (wrong values:
  timing only).  Builds variants/libmodfx_lmexp<bits>.so; run one with
  MODFX_LIB=$PWD/variants/libmodfx_lmexp5.so python scripts/quick_logmel_bench.py
    python scripts/logmel_ablation.py 1 2 3 4 5 6 7 19 16"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "mod_extraction_b200/csrc/logmel.cu")).read()


def sub(old, new):
    global src
    assert src.count(old) == 1, old
    src = src.replace(old, new)


sub("                const int o = moff[m], cnt = moff[m + 1] - o;",
    "#if (MODFX_EXP & 1)\n                const int o = moff[m], cnt = min(1, (int)moff[m + 1] - o);\n#else\n"
    "                const int o = moff[m], cnt = moff[m + 1] - o;\n#endif")
sub("        const bool active = t0 + lf < a.n_frames;", "        const bool active = (t0 + lf < a.n_frames) && !(MODFX_EXP & 2);")
sub("                    *op = *ep;", "#if (MODFX_EXP & 4)\n                    if (a.n_frames < 0) *op = *ep;\n#else\n                    *op = *ep;\n#endif")
sub("            if (t < live) {", """#if (MODFX_EXP & 16)
            {   // store-pattern probe: item `it` writes bands [64 (it & 3), +64) x frames [32 (it >> 2), +32): 128-byte segments
                const int q = it >> 2, sub_ = it & 3;
                const int fr = q * 32 + (tid & 31);
                if (fr < a.n_frames) {
#pragma unroll 8
                    for (int i = 0; i < 16; ++i) {
                        const int m = sub_ * 64 + (tid >> 5) * 16 + i;
                        orow[(int64_t)m * a.n_frames + fr] = E[kStageOff + i * 128 + tid];
                    }
                }
            }
            if (false) {
#else
            if (t < live) {
#endif""")
os.makedirs("/tmp/lmexp", exist_ok=True)
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
open("/tmp/lmexp/logmel_exp.cu", "w").write(src)
subprocess.check_call([sys.executable, "-m", "mod_extraction_b200._build"], cwd=ROOT, stdout=subprocess.DEVNULL)
objs = [os.path.join(ROOT, "mod_extraction_b200/lib/obj", f) for f in os.listdir(os.path.join(ROOT, "mod_extraction_b200/lib/obj"))
        if f.endswith(".o") and f != "logmel.o"]
arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
for bits in sys.argv[1:]:
    o = f"/tmp/lmexp/logmel_{bits}.o"
    subprocess.check_call(["nvcc", *arch, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "177",
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "mod_extraction_b200/csrc"),
                           f"-DMODFX_EXP={bits}", "-c", "/tmp/lmexp/logmel_exp.cu", "-o", o])
    out = os.path.join(ROOT, f"variants/libmodfx_lmexp{bits}.so")
    subprocess.check_call(["nvcc", "-shared", "-o", out, *objs, o, *arch])
    print("built", out)
