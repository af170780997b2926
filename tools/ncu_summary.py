"""Summarise an .ncu-rep (read here, no GPU needed) into the text kept under profiles/.
    python tools/ncu_summary.py gpurun_out/prof_gemm_tc.ncu-rep > profiles/r1_gemm_tc_ncu.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("smsp__cycles_active.avg", "SMSP active cycles (avg)"),
    ("gpc__cycles_elapsed.max", "elapsed cycles"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) %"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__inst_executed.sum", "instructions"),
    ("sm__sass_inst_executed_op_shared_ld.sum", "LDS"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping / issue"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("no kernels in", path)
        return
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    tens = [h for h in hdr if "tensor" in h and (h.endswith(".sum") or "pct_of_peak" in h)]
    print(f"# {path}: {len(rows) - 2} profiled launches (ncu --set full --clock-control none; replayed, cold caches)")
    for r in rows[2:]:
        print("\n## " + r[col["Kernel Name"]][:160])
        for key, label in KEYS:
            if key in col and r[col[key]] != "":
                print(f"  {label:34s} {r[col[key]]:>18s} {units[col[key]]}")
        for h in tens:
            if h not in dict(KEYS) and r[col[h]] not in ("", "0"):
                print(f"  {h[:60]:60s} {r[col[h]]:>14s} {units[col[h]]}")


if __name__ == "__main__":
    main(sys.argv[1])
