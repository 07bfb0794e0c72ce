"""Turn an .ncu-rep (full set, one kernel) into the short text summary kept under profiles/."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]
STALLS = "smsp__average_warps_issue_stalled_"


def main(rep, out, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, val)}
    for h in list(d):            # some counters carry a section prefix ("TPC.TriageCompute.")
        if h[:1].isupper() and h.count(".") >= 2 and d[h][1] != "":
            d.setdefault(h.split(".", 2)[-1], d[h])
    lines = ["kernel: %s" % d.get("Kernel Name", ("", "?"))[1], "source: %s (ncu --set full --clock-control none, one launch)" % rep]
    for k in KEYS:
        if k in d:
            lines.append("%-95s %s %s" % (k, d[k][1], d[k][0]))
    lines.append("warp stall reasons (cycles per issued instruction):")
    for h in sorted(d):
        if h.startswith(STALLS) and h.endswith("per_issue_active.ratio") and float(d[h][1] or 0) > 0.01:
            lines.append("   %-40s %s" % (h[len(STALLS):-len("_per_issue_active.ratio")], d[h][1]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if traffic_json:
        def b(key):
            u, v = d[key]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        json.dump({"kernel": d["Kernel Name"][1], "dram_bytes_per_launch": b("dram__bytes_read.sum") + b("dram__bytes_write.sum"),
                   "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
                   "source": rep}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
