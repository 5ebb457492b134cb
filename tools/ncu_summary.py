"""Turns gpurun_out/*.ncu-rep + the launch list into the tracked summaries under profiles/ (run in the build container)."""
import csv, io, subprocess, collections, json, os
OUT = "profiles"
os.makedirs(OUT, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic"]
def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d
summary = {}
lines = ["# ncu --set full captures, round 1 (one launch each; `--clock-control none`, cold-cache, serialised)", ""]
for name, rep in [("conv256_pair (gemm_tcgen05_kernel<256,BF16,pair> 256->256 3x3x3 @2x1080x1920)", "gpurun_out/r1_conv256.ncu-rep"),
                  ("conv128_swap (gemm_tcgen05_kernel<256,BF16,swap> 128->128 3x3x3 @2x2160x3840)", "gpurun_out/r1_conv128.ncu-rep"),
                  ("swiglu_pair (gemm_tcgen05_kernel<256,SWIGLU,pair> 97200x13824x2560)", "gpurun_out/r1_swiglu.ncu-rep"),
                  ("attn_varlen_kernel (243 windows x 463 tokens x 20 heads; software-pipelined v3)", "gpurun_out/r1_attn_v3.ncu-rep"),
                  ("upsample_shuffle (gemm_tcgen05_kernel<256,BF16,pair> + pixel-shuffle store, 256 ch 2x1080x1920 -> 2x2160x3840)", "gpurun_out/r1_upsample_v2.ncu-rep"),
                  ("conv_shortcut_1x1x1 (gemm_tcgen05_kernel<256,BF16,swap> 256->128 @2x2160x3840)", "gpurun_out/r1_shortcut.ncu-rep"),
                  ("groupnorm_apply_kernel (2x2160x3840x128)", "gpurun_out/prof_gnapply_r1.ncu-rep"),
                  ("groupnorm_stats_kernel (2x2160x3840x128)", "gpurun_out/prof_gnstats_r1.ncu-rep")]:
    if not os.path.exists(rep):
        continue
    d = raw(rep)
    lines.append(f"## {name}")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    rec = {}
    for k in KEYS:
        if k in d:
            lines.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
            rec[k] = d[k][0]
    summary[name.split(" ")[0]] = rec
    lines.append("")
open(f"{OUT}/ncu_full_r1.md", "w").write("\n".join(lines))
json.dump(summary, open(f"{OUT}/ncu_full_r1.json", "w"), indent=1)
# launch list
txt = open("gpurun_out/launches_r1.csv").read().splitlines()
start = [i for i, l in enumerate(txt) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(io.StringIO("\n".join(txt[start:]))))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = r["Kernel Name"]
    short = n.split("(")[0].replace("void ", "").strip()
    if "at::" in short or "elementwise" in short or "Copy" in short or "cub::" in short or "distribution" in short:
        short = "torch (weight/input synthesis, glue: cat/clamp/pad/normalise/to)"
    agg[short][0] += 1
    agg[short][1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
out = ["# ncu launch list, round 1: `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --workload 720p --steps 1 --warmup 1`",
       "", f"{len(rows)} launches (engine build + warm-up step + timed step + e2e step), total {tot / 1e6:.1f} ms device time (cold-cache, serialised: compare shares).", "",
       "| kernel | launches | ms | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {v[0]} | {v[1] / 1e6:.2f} | {100 * v[1] / tot:.1f}% |")
open(f"{OUT}/launches_r1.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))
