#!/usr/bin/env python3
"""Turn an .ncu-rep capture (ncu --set full --import-source on, ONE kernel launch) into the three text extracts that
are committed under profiles/:  <prefix>_ncu_selected.json, <prefix>_ncu_details.csv, <prefix>_sass_regions.txt.

usage: tools/ncu_extract.py gpurun_out/prof_x.ncu-rep profiles/r1n_trace_closest_c2 ["title line"]
"""
import csv, io, json, subprocess, sys
from collections import Counter

SELECTED = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], check=True, capture_output=True, text=True).stdout


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep

    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    names, units, vals = raw[0], raw[1], raw[2]
    sel = {}
    for key in ["Kernel Name", "Grid Size", "Block Size"] + SELECTED:
        if key in names:
            i = names.index(key)
            sel[key] = {"value": vals[i], "unit": units[i]}
    json.dump(sel, open(prefix + "_ncu_selected.json", "w"), indent=1)
    open(prefix + "_ncu_details.csv", "w").write(ncu(rep, "details"))

    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    hdr_i = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    hdr, rows = src[hdr_i], [r for r in src[hdr_i + 1:] if len(r) > 6]
    c_src, c_smp = hdr.index("Source"), hdr.index("# Samples")
    c_exec, c_thr = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    ins = [(r[c_src].strip(), int(r[c_exec]), int(r[c_thr]), int(r[c_smp])) for r in rows]
    total = sum(i[1] for i in ins)
    samples = sum(i[3] for i in ins) or 1
    out = ["# %s" % title,
           "# total warp instructions %.1f M, stall samples %d; regions = runs of instructions with equal execution count" % (total / 1e6, samples),
           "# [first-last] n_instr  exec(M warp-inst each)  total(M)  avg active lanes  % of stall samples  opcode mix"]
    a = 0
    while a < len(ins):
        b = a
        while b + 1 < len(ins) and ins[b + 1][1] == ins[a][1]:
            b += 1
        seg = ins[a:b + 1]
        tot = sum(i[1] for i in seg)
        if tot >= 0.002 * total:
            lanes = sum(i[2] for i in seg) / max(tot, 1)
            ops = Counter(i[0].split()[1].split(".")[0] if i[0].startswith("@") else i[0].split()[0].split(".")[0] for i in seg)
            out.append("[%3d-%3d] n=%3d exec=%7.2fM total=%7.1fM (%4.1f%%) lanes=%4.1f stalls=%5.2f%%  %s" % (
                a, b, len(seg), seg[0][1] / 1e6, tot / 1e6, 100.0 * tot / total, lanes,
                100.0 * sum(i[3] for i in seg) / samples, dict(ops.most_common(6))))
        a = b + 1
    open(prefix + "_sass_regions.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    print(json.dumps({k: v["value"] for k, v in sel.items()}, indent=1))


if __name__ == "__main__":
    main()
