"""profiles/r1_ncu_summary_f32_umma.md from the raw ncu CSV export of the tcgen05 (UMMA) float32 kernels."""
import csv, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU ex2) pipe % of peak"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe (tcgen05.ld/st) %"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (realtime)"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("smsp__inst_executed.sum", "warp instructions")]
rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r1_ncu_full_f32_umma_raw.csv"))))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
names = ["Kuf forward (rowsum mode 0)", "Kdiag forward (rowsum mode 1)", "Kuf backward, dZ part (kuf_dz_kernel)",
         "Kuf backward, dW pass (rowsum mode 3)", "Kdiag backward (rowsum mode 2)"]
out = ["# ncu --set full, round 1: the float32 tcgen05 (UMMA) kernels of one hot-path step at config 3\n",
       "Command (after a plain run of the same command exited 0):",
       "`ncu --set full --clock-control none --import-source on -k regex:\"rowsum_kernel|kuf_dz_kernel\" -s 5 -c 5 python tools/run_step.py --steps 3 --dtype f32`",
       "Raw export: `r1_ncu_full_f32_umma_raw.csv`; launch list of the bench command: `r1_launches_bench_f32_umma.csv`.",
       "SASS of these kernels shows `UTCHMMA` (tcgen05.mma), `LDTM` / `STTM` (tcgen05.ld / st), `LDGSTS` (cp.async), `SYNCS` (mbarrier).",
       "ncu times are cold-cache and serialised: compare shares, not absolutes, with `bench.py`.\n"]
for r, label in zip(rows[2:], names):
    out.append("## %s -- `%s`\n" % (label, r[idx["Kernel Name"]]))
    out.append("| metric | value |\n|---|---|")
    for k, lab in WANT:
        if k in idx:
            out.append("| %s | %s %s |" % (lab, r[idx[k]], units[idx[k]]))
    st = [(float(r[i]) if r[i] else 0.0, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
          for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    st.sort(reverse=True)
    out.append("| top stall reasons (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))
    out.append("")
open(os.path.join(ROOT, "profiles", "r1_ncu_summary_f32_umma.md"), "w").write("\n".join(out) + "\n")
print("written")
