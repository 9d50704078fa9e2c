"""Summarise ncu outputs into profiles/ (launch-list shares and selected raw metrics)."""
import csv, sys
from collections import defaultdict

def launches(path, header):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0][:90]
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    out = list(header)
    out.append("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        out.append(f"{v[1]:10.3f} ms {100*v[1]/tot:5.1f}%  x{v[0]:3d}  {k}")
    out.append(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    return "\n".join(out)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]

def raw(path_csv):
    rows = list(csv.reader(open(path_csv)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append("--- " + r[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"{w:80s} {units[i]:14s} {r[i]}")
    return "\n".join(out)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2], sys.argv[3:]))
    else:
        print(raw(sys.argv[2]))
