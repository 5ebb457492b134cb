#!/bin/bash
# One GPU-box session of round 2: everything is logged under gpurun_out/ (merged back by gpurun).
# usage (on the box, from the repo root): bash tools/gpu_session.sh [stage ...]   (default: all stages)
set -u
mkdir -p gpurun_out
STAGES="${*:-tests glb detail ncu_attn bench}"
for s in $STAGES; do
  echo "=== stage $s $(date +%T)"
  case $s in
    tests)    timeout 900 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/r2_pytest.log 2>&1; tail -45 gpurun_out/r2_pytest.log ;;
    glb)      timeout 600 python tools/gpu_library_baseline.py --ops --phases cfg2,4k_shard --out gpurun_out/r2_gpu_library_baseline.json > gpurun_out/r2_glb.log 2>&1; tail -3 gpurun_out/r2_glb.log ;;
    detail)   timeout 600 python bench.py --steps 3 --warmup 2 --phases --detail --no-cpu-baseline --lib-baseline none > gpurun_out/r2_bench_4k_detail.json 2> gpurun_out/r2_bench_4k_detail.txt; head -40 gpurun_out/r2_bench_4k_detail.txt ;;
    ncu_attn) timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_varlen -c 1 -f -o gpurun_out/r2_attn python tools/perf_conv_one.py attn > gpurun_out/r2_ncu_attn.log 2>&1; tail -3 gpurun_out/r2_ncu_attn.log ;;
    bench)    timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_4k.json 2> gpurun_out/r2_bench_4k.err; cut -c1-1500 gpurun_out/r2_bench_4k.json ;;
    bench1080) timeout 600 python bench.py --workload 1080p --steps 5 --warmup 3 --phases --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_1080p.json 2> gpurun_out/r2_bench_1080p_phases.txt; cut -c1-600 gpurun_out/r2_bench_1080p.json ;;
    glb_ops)  timeout 600 python tools/gpu_library_baseline.py --ops --out gpurun_out/r2_gpu_library_ops.json > gpurun_out/r2_glb_ops.log 2>&1; grep -A8 '"attn_243' gpurun_out/r2_glb_ops.log | head -12 ;;
    bench7b)  timeout 600 python bench.py --workload 4k_shard_7b --steps 3 --warmup 2 --phases --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_4k_shard_7b.json 2> gpurun_out/r2_bench_4k_shard_7b_phases.txt; cut -c1-500 gpurun_out/r2_bench_4k_shard_7b.json ;;
    clip64)   timeout 900 python bench.py --workload 4k_clip64 --steps 2 --warmup 1 --phases --lib-baseline none --no-cpu-baseline --no_graph > gpurun_out/r2_bench_4k_clip64.json 2> gpurun_out/r2_bench_4k_clip64_phases.txt; cut -c1-500 gpurun_out/r2_bench_4k_clip64.json ;;
    sweep)    for T in 16 32 64 128; do timeout 600 python bench.py --workload vae_decode_T$T --steps 3 --warmup 2 --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_vae_decode_T$T.json 2> gpurun_out/r2_bench_vae_decode_T$T.err; cut -c1-330 gpurun_out/r2_bench_vae_decode_T$T.json; echo; done ;;
    sweep2)   for T in 16 128; do timeout 600 python bench.py --workload vae_decode_T$T --steps 2 --warmup 1 --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_vae_decode_T$T.json 2> gpurun_out/r2_bench_vae_decode_T$T.err; cut -c1-330 gpurun_out/r2_bench_vae_decode_T$T.json; echo; done ;;
    ab_wr)    for cfg in "SVR2_CONV_WR=0" "SVR2_CONV_WR=2 SVR2_CONV_WR_BO=0" "SVR2_CONV_WR=2 SVR2_CONV_WR_BO=1"; do
                echo "--- $cfg"; env $cfg timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "conv3d" 2>&1 | tail -4
                for i in 1 2 3; do env $cfg python tools/perf_conv_one.py conv128 2>&1 | tail -1; done
              done 2>&1 | tee gpurun_out/r2_ab_wr.log
              for v in 0 1; do echo "--- SVR2_ATTN_POLY=$v"; SVR2_ATTN_POLY=$v timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "attn" 2>&1 | tail -2
                for i in 1 2 3; do SVR2_ATTN_POLY=$v python tools/perf_conv_one.py attn 2>&1 | tail -1; done
              done 2>&1 | tee gpurun_out/r2_ab_attn.log ;;
    ncu_wr)   timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wreuse -c 1 -f -o gpurun_out/r2_conv_wr python tools/perf_conv_one.py conv128 > gpurun_out/r2_ncu_conv_wr.log 2>&1; tail -2 gpurun_out/r2_ncu_conv_wr.log
              SVR2_CONV_WR=0 timeout 600 ncu --set full --clock-control none -k regex:gemm_tcgen05 -c 1 -f -o gpurun_out/r2_conv128_generic python tools/perf_conv_one.py conv128 > gpurun_out/r2_ncu_conv128_generic.log 2>&1; tail -2 gpurun_out/r2_ncu_conv128_generic.log ;;
    ab_wr2)   for cfg in "SVR2_CONV_WR=0" "SVR2_CONV_WR=1"; do echo "--- $cfg"
                for c in conv128 conv_sc; do for i in 1 2; do env $cfg PERF_REPS=10 python tools/perf_conv_one.py $c 2>&1 | tail -1; done; done
              done 2>&1 | tee gpurun_out/r2_ab_wr2.log
              timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "conv3d" 2>&1 | tail -3 ;;
    ab_attn2) for v in 0 1; do echo "--- SVR2_ATTN_PTMEM=$v"; SVR2_ATTN_PTMEM=$v timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -k "attn or attention or dit_vs_golden" 2>&1 | tail -3
                for i in 1 2 3; do SVR2_ATTN_PTMEM=$v PERF_REPS=20 python tools/perf_conv_one.py attn 2>&1 | tail -1; done
              done 2>&1 | tee gpurun_out/r2_ab_attn2.log ;;
    ab_wr3)   for cfg in "SVR2_CONV_WR_RING=35" "SVR2_CONV_WR_RING=43"; do echo "--- $cfg"
                for c in conv128 conv_sc; do for i in 1 2; do env $cfg PERF_REPS=10 python tools/perf_conv_one.py $c 2>&1 | tail -1; done; done
              done 2>&1 | tee gpurun_out/r2_ab_wr3.log
              SVR2_CONV_WR_RING=43 timeout 200 python -m pytest tests/test_ops_gpu.py -q -x -k "conv3d" 2>&1 | tail -2 ;;
    ab_wrp)   timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "conv3d" 2>&1 | tail -4
              for cfg in "SVR2_CONV_WR_PAIR=0" "SVR2_CONV_WR_PAIR=1"; do echo "--- $cfg"
                for i in 1 2 3; do env $cfg PERF_REPS=10 python tools/perf_conv_one.py conv256 2>&1 | tail -1; done
              done 2>&1 | tee gpurun_out/r2_ab_wrp.log
              timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 1 -f -o gpurun_out/r2_conv256_wr python tools/perf_conv_one.py conv256 > gpurun_out/r2_ncu_conv256_wr.log 2>&1; tail -2 gpurun_out/r2_ncu_conv256_wr.log ;;
    ab_noshift) for cfg in "SVR2_WR_NOSHIFT=0" "SVR2_WR_NOSHIFT=1"; do echo "--- $cfg"
                for c in conv256 conv128; do env $cfg PERF_REPS=10 python tools/perf_conv_one.py $c 2>&1 | tail -1; done
                env $cfg timeout 300 ncu --metrics sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed.sum --clock-control none -k regex:"gemm_tcgen05|conv_wreuse" -c 1 python tools/perf_conv_one.py conv256 2>&1 | grep -E "tensor_cycles|per_second|duration" 
                env $cfg timeout 300 ncu --metrics sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum --clock-control none -k regex:"gemm_tcgen05|conv_wreuse" -c 1 python tools/perf_conv_one.py conv128 2>&1 | grep -E "tensor_cycles|per_second|duration"
              done 2>&1 | tee gpurun_out/r2_ab_noshift.log ;;
    ab_wr4)   timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "conv3d" 2>&1 | tail -2
              for cfg in "SVR2_CONV_WR=0" "SVR2_CONV_WR=1"; do echo "--- $cfg"
                for c in conv256 conv128 conv_sc; do for i in 1 2; do env $cfg PERF_REPS=10 python tools/perf_conv_one.py $c 2>&1 | tail -1; done; done
                for c in conv256 conv128; do env $cfg timeout 300 ncu --metrics sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum --clock-control none -k regex:"gemm_tcgen05|conv_wreuse" -c 1 python tools/perf_conv_one.py $c 2>&1 | grep -E "tensor_cycles|per_second|duration"; done
              done 2>&1 | tee gpurun_out/r2_ab_wr4.log ;;
    ab_bench) for v in 0 1; do SVR2_CONV_WR_PAIR=$v timeout 600 python bench.py --steps 3 --warmup 2 --lib-baseline none --no-cpu-baseline --no_graph --phases > gpurun_out/r2_ab_bench_pair$v.json 2> gpurun_out/r2_ab_bench_pair$v.txt; cut -c1-200 gpurun_out/r2_ab_bench_pair$v.json; echo; done ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --workload 1080p --steps 1 --warmup 1 --no_graph --lib-baseline none --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1; tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300 ;;
    ncu_conv) timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 1 -f -o gpurun_out/r2_conv256 python tools/perf_conv_one.py conv256 > gpurun_out/r2_ncu_conv256.log 2>&1; tail -2 gpurun_out/r2_ncu_conv256.log ;;
    debug_native) timeout 300 python tools/debug_native.py > gpurun_out/r2_debug_native.log 2>&1; cat gpurun_out/r2_debug_native.log | tail -14 ;;
    tests_attn) timeout 600 python -m pytest tests -m gpu -q -k "attn or attention or dit_vs_golden or native" > gpurun_out/r2_pytest_attn.log 2>&1; tail -8 gpurun_out/r2_pytest_attn.log ;;
    ab_upsample) for v in "" NO_RUNIF NO_ROWSCALE NO_BOTH; do for i in 1 2 3; do if [ -z "$v" ]; then L=""; else L="$PWD/comfyui-seedvr2_videoupscaler_b200/csrc/libsvr2_ab_$v.so"; fi; echo -n "lib=${v:-default} "; SVR2_LIB=$L python tools/perf_conv_one.py upsample 2>&1 | tail -1; done; done | tee gpurun_out/r2_ab_upsample.log ;;
    mgpu8)    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
              timeout 900 $TR bench.py --gpus 8 --workload 4k_shard_7b --steps 3 --warmup 2 --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_4k_shard_7b_n8.json 2> gpurun_out/r2_bench_4k_shard_7b_n8.err; cut -c1-400 gpurun_out/r2_bench_4k_shard_7b_n8.json
              for T in 16 128; do timeout 900 $TR bench.py --gpus 8 --workload vae_decode_T$T --steps 2 --warmup 1 --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_vae_decode_T${T}_n8.json 2> gpurun_out/r2_bench_vae_decode_T${T}_n8.err; cut -c1-330 gpurun_out/r2_bench_vae_decode_T${T}_n8.json; echo; done ;;
    mgpu8b)   TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
              timeout 600 $TR bench.py --gpus 8 --workload 4k_shard_7b --steps 2 --warmup 1 --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_4k_shard_7b_n8.json 2> gpurun_out/r2_bench_4k_shard_7b_n8.err; cut -c1-400 gpurun_out/r2_bench_4k_shard_7b_n8.json; echo
              timeout 600 $TR bench.py --gpus 8 --workload vae_decode_T128 --steps 1 --warmup 1 --lib-baseline none --no-cpu-baseline > gpurun_out/r2_bench_vae_decode_T128_n8.json 2> gpurun_out/r2_bench_vae_decode_T128_n8.err; cut -c1-330 gpurun_out/r2_bench_vae_decode_T128_n8.json; echo ;;
    perf_up)  timeout 300 python tools/perf_upsample.py 2>&1 | tee gpurun_out/r2_perf_upsample.log ;;
    *) echo "unknown stage $s" ;;
  esac
done
echo "=== done $(date +%T)"
