"""Summarise an .ncu-rep (raw page) into a markdown table of the metrics the roofline discussion uses."""
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'local_load_bytes', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']


def main(rep, title):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none)\n")
    for r in rows[2:]:
        print(f"## {r[ki][:110]}\n\n| metric | value | unit |\n|---|---|---|")
        vals = {}
        for h in KEEP:
            if h in hdr:
                v = r[hdr.index(h)]
                vals[h] = v
                print(f"| {h} | {v} | {units[hdr.index(h)]} |")
        try:
            f = lambda k: float(vals[k].replace(",", ""))
            per_cycle = (f('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed')
                         + f('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed')
                         + 2 * f('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed'))
            peak = 2 * f('sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained')
            t = f('gpu__time_duration.sum') * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}[units[hdr.index('gpu__time_duration.sum')]]
            cyc = f('sm__cycles_elapsed.avg')
            print(f"| executed fp64 flop/cycle (dadd+dmul+2*dfma) | {per_cycle:.1f} of {peak:.0f} = {per_cycle / peak:.3f} | flop/cycle |")
            print(f"| executed fp64 flop per launch | {per_cycle * cyc:.4e} | flop |")
            print(f"| executed fp64 rate | {per_cycle * cyc / t / 1e12:.2f} | TFLOP/s |")
        except Exception as exc:
            print(f"| flop summary | unavailable ({exc}) | |")
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
