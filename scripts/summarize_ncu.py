"""Turn an .ncu-rep into a small text summary (key raw metrics + stall mix + hottest SASS lines).
Usage: python scripts/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path):
    raw = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    print(f"# {path}")
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        print(f"\n## {name[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:70s} {row[i]:>18s} {units[i]}")
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    # the source page repeats a 2-line header per kernel; handle the first kernel only unless split is easy
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for blk in blocks:
        rows = blk["rows"]
        if not rows:
            continue
        h = rows[0]
        if "Source" not in h:
            continue
        isrc, ie, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = {c: 0 for _, c in stall}
        data = []
        for r in rows[1:]:
            if len(r) <= ie:
                continue
            try:
                data.append((r[isrc], int(r[ie] or 0), int(r[ismp] or 0)))
            except ValueError:
                continue
            for i, c in stall:
                try:
                    tot[c] += int(r[i] or 0)
                except ValueError:
                    pass
        s = sum(tot.values()) or 1
        ti = sum(d[1] for d in data) or 1
        ts = sum(d[2] for d in data) or 1
        print(f"\n### stall mix of {blk['name'][:90]}")
        print("  " + ", ".join(f"{c[6:]} {v / s * 100:.1f}%" for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
        print(f"### hottest SASS by stall samples (share of samples, share of executed instructions); {len(data)} SASS lines")
        for d in sorted(data, key=lambda d: -d[2])[:12]:
            print(f"  {d[2] / ts * 100:5.2f}%  {d[1] / ti * 100:5.2f}%  {d[0][:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
