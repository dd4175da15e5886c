#!/usr/bin/env python3
"""Turns `ncu --set full` captures of the closest-hit launch on C2 and on C4 into the two small JSON files bench.py reads for
`roofline.traffic` (DRAM bytes per launch cannot be measured live: a number printed under a profiler is never a bench value).

  tools/capture_dram.py <c2.ncu-rep> <c4.ncu-rep> <git sha of the build that was profiled>
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout)))
    names, units, vals = rows[0], rows[1], rows[2]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return int(round(f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]))


def main():
    c2, c4, sha = sys.argv[1], sys.argv[2], sys.argv[3]
    for rep, name, what in ((c2, "trace_closest_c2_dram.json", "C2, 8M incoherent rays"), (c4, "trace_closest_c4_dram.json", "C4, 10M triangles, 8M incoherent rays")):
        if not os.path.exists(rep):
            continue
        m = raw(rep)
        rd, wr = to_bytes(*m["dram__bytes_read.sum"]), to_bytes(*m["dram__bytes_write.sum"])
        out = {"kernel": m["Kernel Name"][0] + " (" + what + ")", "head": sha, "capture": os.path.basename(rep) + " (ncu --set full --clock-control none)",
               "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
               "l2_hit_rate": float(m["lts__t_sector_hit_rate.pct"][0]), "l1_hit_rate": float(m["l1tex__t_sector_hit_rate.pct"][0]),
               "duration_ms_under_ncu": float(m["gpu__time_duration.sum"][0]) * (1e-3 if m["gpu__time_duration.sum"][1] in ("us", "usecond") else 1.0)}
        json.dump(out, open(os.path.join(ROOT, "profiles", name), "w"), indent=1)
        print(name, out)


if __name__ == "__main__":
    main()
