"""Summarise an ncu report (ncu --set full) as a small text table: one column per captured
launch, the metrics DESIGN.md and bench.py quote.  Usage: ncu_summary.py REPORT.ncu-rep [OUT.txt]
Also prints the per-kernel DRAM traffic as JSON (for profiles/pileup_traffic.json)."""
import csv
import io
import json
import subprocess
import sys

METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
sm__throughput.avg.pct_of_peak_sustained_elapsed smsp__issue_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__grid_size smsp__inst_executed.sum lts__t_sector_hit_rate.pct
l1tex__t_sector_hit_rate.pct l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__thread_inst_executed_per_inst_executed.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio""".split()


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [r[ki].replace("<unnamed>::", "").split("(")[0] for r in data]
    lines = ["%-78s %s" % ("Kernel Name", " | ".join(names))]
    traffic = {}
    for m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        lines.append("%-78s %s" % ("%s [%s]" % (m, units[i]), " | ".join(r[i] for r in data)))
    for r, n in zip(data, names):
        def val(m):
            i = hdr.index(m)
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            return float(r[i].replace(",", "")) * scale
        traffic[n] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
