"""Rebuilds the round's profile summaries from the raw ncu outputs brought back in gpurun_out/ (scratch):
  gpurun_out/prof_tc.ncu-rep  (ncu --set full of one tc_chunk_kernel launch)  -> r01_tc_kernel_ncu_summary.md, ncu_traffic.json
  gpurun_out/launches.csv     (ncu --metrics gpu__time_duration.sum launch list) -> r01_launch_list.md
Run here (no GPU needed):  python profiles/make_summaries.py"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", "prof_tc.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
rows = [r for r in rows if len(r) > 10]
m = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed", "smsp__mem_tensor_reads_op_utcmma_matrix_c.sum.pct_of_peak_sustained_elapsed",
        "smsp__mem_tensor_writes_op_utcmma.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum"]
def num(k):
    u, v = m[k]
    return float(v.replace(",", "")), u
lines = ["| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in m:
        lines.append("| %s | %s | %s |" % (k, m[k][1], m[k][0]))
rd, ru = num("dram__bytes_read.sum"); wr, wu = num("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd *= scale[ru]; wr *= scale[wu]
n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
alg = n_rows * 260
t, tu = num("gpu__time_duration.sum")
traffic = {"kernel": "tc_chunk_kernel<true,false>", "config": "bench.py default: %d x 64 fp32, k=256, one launch" % n_rows,
           "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
           "algorithmic_bytes_per_launch": alg, "ratio": (rd + wr) / alg,
           "gpu_time_under_ncu": "%s %s" % (m["gpu__time_duration.sum"][1], tu),
           "how": "ncu --set full --clock-control none --import-source on -k regex:tc_chunk_kernel -s 3 -c 1 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e (r01)"}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", "_ncu_table.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)); print(json.dumps(traffic))

lst = os.path.join(ROOT, "gpurun_out", "launches.csv")
rows = [r for r in csv.reader(open(lst)) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    k = r[ik][:90]; v = float(r[iv].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
L = ["| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda t: -t[1][1])[:16]:
    L.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))
ours = {k: a for k, a in agg.items() if "bkm::" in k}
step = sum(a[1] / a[0] for a in ours.values())
L.append("")
L.append("Our kernels, one launch each per Lloyd step (averages above): total %.1f us; `tc_chunk_kernel` share %.1f %%."
         % (step / 1e3, 100 * max(a[1] / a[0] for a in ours.values()) / step))
open(os.path.join(ROOT, "profiles", "_launch_table.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L))
