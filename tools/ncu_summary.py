"""Summarise an .ncu-rep (ncu --set full) into a few lines per kernel: python tools/ncu_summary.py rep.ncu-rep [out.txt]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
    "local_load", "local_store", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_barriers",
]


def main(rep, out=None, traffic_json=None, compute_json=None):
    """traffic_json: also write {kernel base name: {"dram_bytes": read+write, "time_us": ...}} for the longest launch
    of every kernel (bench.py's roofline.traffic reads it).  compute_json: the issue-slot / fp64-pipe ceilings of every
    kernel (bench.py's roofline_compute): warp instructions / (SMs x 4 schedulers x elapsed cycles)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    traffic = {}
    compute = {}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tmul = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        base = name.replace("void ", "").split("(")[0].split("<")[0].split("::")[-1].replace("_kernel", "")
        i_t, i_r, i_w = hdr.index("gpu__time_duration.sum"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        t_us = float(r[i_t]) * tmul.get(units[i_t], 1.0)
        byt = float(r[i_r]) * mult.get(units[i_r], 1.0) + float(r[i_w]) * mult.get(units[i_w], 1.0)
        if base not in traffic or t_us > traffic[base]["time_us"]:
            traffic[base] = {"dram_bytes": byt, "time_us": t_us, "kernel": name[:80]}

            def col(metric):
                for i, hname in enumerate(hdr):
                    if hname == metric:
                        try:
                            return float(r[i].replace(",", ""))
                        except ValueError:
                            return None
                return None
            inst, cyc = col("smsp__inst_executed.sum"), col("sm__cycles_elapsed.max")
            if inst and cyc:
                sms = 148
                compute[base] = {
                    "bound": "issue slots", "warp_instructions": inst, "elapsed_cycles": cyc, "sms": sms,
                    "achieved": inst / cyc / (sms * 4), "peak": 1.0, "unit": "warp instructions / scheduler / cycle",
                    "frac": inst / cyc / (sms * 4),
                    "issue_active_pct_of_active_cycles": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "fp64_pipe_active_pct": col("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                    "warps_active_pct": col("sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "registers_per_thread": col("launch__registers_per_thread"), "time_us": t_us, "kernel": name[:80]}
        lines.append(f"== {r[hdr.index('Kernel Name')][:70]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w or h.endswith("." + w) or (w in ("local_load", "local_store") and w in h and h.endswith(".sum")):
                    lines.append(f"   {h} = {r[i]} {units[i]}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")
    if traffic_json:
        import json
        json.dump(traffic, open(traffic_json, "w"), indent=1)
    if compute_json:
        import json
        json.dump(compute, open(compute_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
