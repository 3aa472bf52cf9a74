"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of counters the roofline
argument rests on.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_elapsed.avg.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        print(f"kernel: {name}")
        d = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k][0]:>20s} {d[k][1]}")
        print("  -- warp stall reasons (warps per issue-active cycle) --")
        st = [(float(v[0]), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v[0]]
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"  {h:75s} {v:20.3f}")
        try:
            rd, wr = float(d["dram__bytes_read.sum"][0]), float(d["dram__bytes_write.sum"][0])
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * scale[d["dram__bytes_read.sum"][1]] + wr * scale[d["dram__bytes_write.sum"][1]]
            ms = float(d["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[d["gpu__time_duration.sum"][1]]
            print(f"  dram traffic per launch: {tot / 1e9:.4f} GB  ->  {tot / ms / 1e6:.1f} GB/s under the profiler")
        except Exception as e:  # noqa: BLE001
            print("  (traffic summary unavailable:", e, ")")


if __name__ == "__main__":
    main()
