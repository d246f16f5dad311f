"""Reads an .ncu-rep (ncu -i ... --page raw --csv) and writes the counters the docs cite, one block per launch."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def main(rep, out_txt, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines, traffic = [], []
    for r in data:
        name = r[hdr.index("Kernel Name")]
        lines.append("== launch %s  %s" % (r[0], name))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("   %-80s %s %s" % (k, r[i], units[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd *= mult.get(units[hdr.index("dram__bytes_read.sum")], 1)
            wr *= mult.get(units[hdr.index("dram__bytes_write.sum")], 1)
            traffic.append(rd + wr)
        except (ValueError, IndexError):
            pass
    open(out_txt, "w").write("\n".join(lines) + "\n")
    if traffic_json and traffic:
        json.dump({"source": rep, "dram_bytes_per_launch": sum(traffic) / len(traffic),
                   "per_launch": traffic, "note": "dram__bytes_read.sum + dram__bytes_write.sum of the captured "
                   "k_synth_pass launches (ncu --set full)"}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
