"""profiles/r1_ncu_summary.md + profiles/r1_traffic.json from the raw ncu CSV exports."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe active %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 (scalar) pipe active %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "HMMA (TF32 mma.sync) pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("smsp__inst_executed.sum", "warp instructions")]
SHORT = [("kuf", ", 0,", "kuf_fwd"), ("kuf", ", 1,", "kuf_bwd"), ("kdiag", ", 0>", "kdiag_fwd"), ("kdiag", ", 1>", "kdiag_bwd"),
         ("kuu", ", 0>", "kuu_fwd"), ("kuu", ", 1>", "kuu_bwd")]
out = ["# ncu --set full, round 1: the six kernels of one hot-path step at config 3 (28x28, 5x5, M=750, N=200)\n",
       "Commands (each after a plain run of the same command exited 0):",
       "`ncu --set full --clock-control none --import-source on -k regex:\"mma_kernel|kuu_fast\" -s 6 -c 6 python tools/run_step.py --steps 3`",
       "`ncu --set full --clock-control none --import-source on -k regex:\"tf32_kernel|kuu_fast\" -s 6 -c 6 python tools/run_step.py --steps 3 --dtype f32`",
       "Raw exports: `r1_ncu_full_f64_raw.csv`, `r1_ncu_full_f32_raw.csv`.  DMMA and scalar FP64 share one physical pipe on",
       "B200 (`r1_probe_pipe_calibration.csv`: the two counters sum to ~90-100 % on the mixed probe), so \"FP64 pipe busy\" is",
       "their sum.  ncu times are cold-cache and serialised: compare shares, not absolutes, with `bench.py`.\n"]
traffic = {}
for dt in ("f64", "f32"):
    rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r1_ncu_full_%s_raw.csv" % dt))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic[dt] = {}
    out.append("# %s\n" % ("float64 (FP64 warp-MMA kernels)" if dt == "f64" else "float32 (3xTF32 tensor-core kernels)"))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        out.append("## `%s`\n" % name)
        out.append("| metric | value |\n|---|---|")
        vals = {}
        for k, lab in WANT:
            if k in idx and r[idx[k]] not in ("", "0", "0.000000") or k.startswith(("gpu__time", "launch", "dram")):
                if k in idx:
                    out.append("| %s | %s %s |" % (lab, r[idx[k]], units[idx[k]]))
                    vals[k] = r[idx[k]]
        try:
            busy = float(vals.get("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", 0)) + \
                float(vals.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0))
            if dt == "f64":
                out.append("| **FP64 pipe busy (DMMA + scalar)** | **%.1f %%** |" % busy)
        except Exception:
            pass
        st = [(float(r[i]) if r[i] else 0.0, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
              for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        st.sort(reverse=True)
        out.append("| top stall reasons (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st[:5]))
        out.append("")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tb = float(r[idx["dram__bytes_read.sum"]]) * mult[units[idx["dram__bytes_read.sum"]]] + \
            float(r[idx["dram__bytes_write.sum"]]) * mult[units[idx["dram__bytes_write.sum"]]]
        for a, b, short in SHORT:
            if a in name.split("<")[0] and b in name + ">":
                traffic[dt][short] = tb
open(os.path.join(ROOT, "profiles", "r1_ncu_summary.md"), "w").write("\n".join(out) + "\n")
json.dump({"source": "profiles/r1_ncu_full_{f64,f32}_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, cfg 3)",
           **traffic}, open(os.path.join(ROOT, "profiles", "r1_traffic.json"), "w"), indent=1)
print(json.dumps(traffic))
