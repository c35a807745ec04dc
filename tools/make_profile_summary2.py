"""profiles/<tag>_ncu_summary.md from the raw ncu CSV exports of tools/profile_round.sh (both dtypes, the six kernels
of one hot-path step in launch order).   python tools/make_profile_summary2.py r1z"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1z"
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe active %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 (scalar) pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe % of peak"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe (tcgen05.ld/st) %"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
         "tensor pipe active % (realtime)"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("smsp__inst_executed.sum", "warp instructions")]
LABELS = ["Kuf forward", "Kdiag forward with saved row sums", "Kuu forward", "Kuf backward (dZ, dW, dvariance, dlengthscale)",
          "Kdiag backward from the saved row sums", "Kuu backward"]
KEYS = ["kuf_fwd", "kdiag_fwd", "kuu_fwd", "kuf_bwd", "kdiag_bwd", "kuu_bwd"]
STALL = "smsp__average_warps_issue_stalled_"
out = ["# ncu --set full, %s: the six kernels of one hot-path step at config 3 (28x28, 5x5, M=750, N=200)\n" % TAG,
       "Recipe: `tools/profile_round.sh %s` (each ncu pass after a plain run of the same command exited 0):" % TAG,
       "`ncu --set full --clock-control none --import-source on -k regex:\"mma_kernel|kuu_fast|kdiag_bwd_saved\" -s 6 -c 6 "
       "python tools/run_step.py --steps 3` and the same with `kuf_fwd_kernel|kdiag_fwd_kernel|kuf_bwd_kernel|...` and `--dtype f32`.",
       "Raw exports: `%s_ncu_full_f64_raw.csv`, `%s_ncu_full_f32_raw.csv`; launch lists of the bench command: "
       "`%s_launches_bench_f64.csv`, `%s_launches_bench_f32.csv`; hottest source lines: `%s_hot_*.txt`." % ((TAG,) * 5),
       "DMMA and scalar FP64 share one physical pipe on B200 (`r1_probe_pipe_calibration.csv`), so \"FP64 pipe busy\" is "
       "their sum.  ncu times are cold-cache and serialised: compare shares, not absolutes, with `bench.py`.\n"]
traffic = {"source": "profiles/%s_ncu_full_{f64,f32}_raw.csv (ncu --set full, dram__bytes_read.sum + "
                     "dram__bytes_write.sum per launch, config 3)" % TAG}
for dt, title in (("f64", "float64 (FP64 warp-MMA kernels)"), ("f32", "float32 (tcgen05 / TMEM kernels)")):
    path = os.path.join(ROOT, "profiles", "%s_ncu_full_%s_raw.csv" % (TAG, dt))
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out.append("# %s\n" % title)
    traffic[dt] = {}
    for r, label, key in zip(rows[2:], LABELS, KEYS):
        out.append("## %s -- `%s`\n" % (label, r[idx["Kernel Name"]]))
        out.append("| metric | value |\n|---|---|")
        vals = {}
        for k, lab in WANT:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                v = r[idx[k]]
                try:
                    if float(v.replace(",", "")) == 0.0 and not k.startswith(("gpu__time", "launch", "dram")):
                        continue
                except ValueError:
                    pass
                vals[k] = v
                out.append("| %s | %s %s |" % (lab, v, units[idx[k]]))
        try:
            fp64 = float(vals.get(WANT[4][0], 0)) + float(vals.get(WANT[5][0], 0))
            if fp64 > 0:
                out.append("| **FP64 pipe busy (DMMA + scalar)** | **%.1f %%** |" % fp64)
        except ValueError:
            pass
        stalls = []
        for h, i in idx.items():
            if h.startswith(STALL) and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio"):
                try:
                    stalls.append((float(r[i]), h[len(STALL):].replace(".ratio", "")))
                except ValueError:
                    pass
        if stalls:
            out.append("| top stall reasons (warps per issue) | %s |" %
                       ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]))
        out.append("")
        try:
            def nbytes(k):
                v, u = float(r[idx[k]].replace(",", "")), units[idx[k]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            traffic[dt][key] = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
        except (KeyError, ValueError):
            pass
open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % TAG), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "%s_traffic.json" % TAG), "w"), indent=1)
print("wrote", TAG)
