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
    "tc_chunk_C2": ("prof_tc_r02.ncu-rep", "tc_chunk_kernel<true,false> on C2 10M x 64, k=256; ncu --set full --clock-control none --import-source on -k regex:tc_chunk_kernel --launch-skip 3 -c 1 python tests/shape_bench.py C2 --steps 2"),
    "tc_chunk_C3": ("prof_tc_c3_r02.ncu-rep", "tc_chunk_kernel<true,false> on C3 4,898,431 x 41 (pitch 44), k=100, KDD-shaped cluster sizes; same command with C3"),
    "tc2_assign_C5": ("prof_tc2_c5.ncu-rep", "tc2_assign_kernel on one 15.6M-row chunk of C5 (125M x 128 bf16 per GPU, k=1024, 2 centre slices); ncu --set full --clock-control none -k regex:tc2_assign --launch-skip 6 -c 1 python bench.py --steps 5 --warmup 1 --no-cpu --no-e2e --configs C5"),
    "rowpass_mstep_C5": ("prof_rowpass_c5.ncu-rep", "rowpass_mstep_kernel (label-indexed M-step row pass) on one 15.6M-row chunk of C5; same command with -k regex:rowpass_mstep"),
    "stream_v3_C4": ("prof_stream3.ncu-rep", "stream2_chunk_kernel<7,20,true> with the shared-memory M-step (lane = feature x row half) on C4 15M x 13, k=20; ncu --set full -k regex:stream2_chunk_kernel --launch-skip 3 -c 1 python tests/shape_bench.py C4 --steps 2"),
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

# ---- launch list of `bench.py --steps 2 --warmup 1 --no-cpu` (ncu --metrics gpu__time_duration.sum --clock-control none)
ll = os.path.join(ROOT, "gpurun_out", "launches_r02.csv")
if os.path.exists(ll):
    import collections, re
    rows = [r for r in csv.reader(open(ll)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
        name = re.sub(r"^at::.*?(\w+_kernel|\w+Kernel\w*).*", r"torch: \1", name)
        v = float(r[14].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[13], 1e-3)
        a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
        a[0] += 1; a[1] += v
    ours = {k: v for k, v in agg.items() if k.startswith("bkm::")}
    tot = sum(v[1] for v in ours.values()) or 1.0
    L = ["# Round 2 — launch list of `python bench.py --steps 2 --warmup 1 --no-cpu` under",
         "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600` (first 600 launches: data generation, the C2",
         "warm-up + timed steps, parity check, e2e pass and the start of the sub-configs).  Per-launch times are cold-cache and",
         "serialised: the SHARE of a kernel is what to compare with the bench line, not the absolute.", "",
         "| kernel | launches | total us | share of our kernels | block | grid |", "|---|---|---|---|---|---|"]
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        L.append("| `%s` | %d | %.1f | %.1f %% | %s | %s |" % (k, v[0], v[1], 100 * v[1] / tot, v[2], v[3]))
    other = sum(v[1] for k, v in agg.items() if not k.startswith("bkm::"))
    L += ["", "torch kernels (synthetic data, fills, copies): %d launches, %.1f us in total." %
          (sum(v[0] for k, v in agg.items() if not k.startswith("bkm::")), other)]
    # one C2 Lloyd iteration in launch order: the launches between two finalize_step_fused launches that contain a
    # tc_chunk_kernel<1,0,0> launch longer than 1 ms
    seq = [(re.sub(r"\(.*", "", r[4]).replace("void ", "").strip(),
            float(r[14].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[13], 1e-3)) for r in rows]
    fin = [i for i, (n, _) in enumerate(seq) if n.startswith("bkm::finalize_step_fused")]
    for a_, b_ in zip(fin, fin[1:]):
        win = seq[a_ + 1:b_ + 1]
        if any(n.startswith("bkm::tc_chunk_kernel<1, 0, 0>") and v > 1000 for n, v in win):
            tw = sum(v for _, v in win)
            L += ["", "One C2 (10M x 64, k=256) Lloyd iteration in launch order:", "", "| kernel | us | share of the iteration |", "|---|---|---|"]
            for n, v in win:
                L.append("| `%s` | %.1f | %.1f %% |" % (n[:90], v, 100 * v / tw))
            L.append("| total | %.1f | (bench line: `roofline.kernel_ms` / `ms_per_step` = the fused kernel's share) |" % tw)
            break
    open(os.path.join(ROOT, "profiles", "r02_launch_list.md"), "w").write("\n".join(L) + "\n")
    print("wrote launch list")
