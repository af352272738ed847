"""Per-kernel summary of an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`):
duration, DRAM traffic, pipe utilisation, occupancy, top warp-stall reasons. Writes markdown to stdout."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe inst %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe inst %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe inst %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread-inst"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread-inst"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread-inst"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-inst"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-inst"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-inst"),
    ("smsp__inst_executed.sum", "warp inst executed"),
    ("local_load_bytes", "local load"), ("smsp__inst_executed_op_local_ld.sum", "local ld inst"),
    ("smsp__inst_executed_op_local_st.sum", "local st inst"),
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
             or h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
    for r in rows[2:]:
        print("### %s" % r[col["Kernel Name"]].split("(")[0])
        for k, label in KEYS:
            if k in col and r[col[k]] not in ("", "n/a"):
                print("- %s: %s %s" % (label, r[col[k]], units[col[k]]))
        st = []
        for h in stall:
            try:
                st.append((float(r[col[h]].replace(",", "")), h))
            except ValueError:
                pass
        st.sort(reverse=True)
        print("- top stalls: " + "; ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")
                                                        .replace("_per_issue_active.ratio", "").replace(".ratio", ""), v) for v, h in st[:5]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
