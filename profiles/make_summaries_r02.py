"""Round-2 profile summaries from the raw ncu reports brought back in gpurun_out/ (scratch, untracked):
    python profiles/make_summaries_r02.py            (no GPU needed; ncu imports the .ncu-rep files)
Writes profiles/r02_ncu_<name>.md for every report listed below that exists."""
import csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORTS = {
    "stream_v1_C4": ("prof_stream.ncu-rep", "stream_chunk_kernel<7,10,true> (one row per thread) on C4 15M x 13, k=20; ncu --set full -k regex:stream_chunk -s 4 -c 1 python tests/shape_bench.py C4 --steps 2"),
    "stream_v2_C4": ("prof_stream2.ncu-rep", "stream2_chunk_kernel<7,20,true> (two rows per thread) on C4 15M x 13, k=20; same command"),
    "tc2_assign_C5s": ("prof_tc2.ncu-rep", "tc2_assign_kernel on 8M x 128 bf16, k=1024 (slice of C5); ncu --set full -k regex:tc2_assign -s 4 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --configs C5s"),
    "rowpass_mstep_v0": ("prof_rowpass2.ncu-rep", "rowpass_mstep_kernel (early per-warp label-scan version, 4M x 128 bf16, k=1024) — the profile that showed the gather loads were not in flight together"),
    "tc_chunk_C2": ("prof_tc_r02.ncu-rep", "tc_chunk_kernel<true,false> on C2 10M x 64, k=256; ncu --set full -k regex:tc_chunk_kernel -s 3 -c 1 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-configs"),
}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"
for name, (fn, what) in REPORTS.items():
    rep = os.path.join(ROOT, "gpurun_out", fn)
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if len(r) > 10]
    m = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    L = ["# Round 2 — ncu `--set full --clock-control none` capture: %s" % name, "", what, "",
         "Numbers under a profiler are not bench values (cold caches, serialised launches); the bench lines are in `r02_bench_*.json`.",
         "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            L.append("| %s | %s | %s |" % (k, m[k][1], m[k][0]))
    st = sorted(((float(v[1].replace(",", "")), k[len(STALLS):]) for k, v in m.items()
                 if k.startswith(STALLS) and "not_issued" not in k), reverse=True)
    tot = sum(s for s, _ in st) or 1.0
    L += ["", "Warp stall samples (all): " + ", ".join("%s %.0f%%" % (k, 100 * s / tot) for s, k in st[:8])]
    open(os.path.join(ROOT, "profiles", "r02_ncu_%s.md" % name), "w").write("\n".join(L) + "\n")
    print("wrote", name)
