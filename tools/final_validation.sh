#!/usr/bin/env bash
# The final validation pass of a round, on a GPU box (one B200): full GPU test suite with error recording, the bench lines
# (inference with torch_cuda / cpu_baseline / variants, train, Qwen), VAE decode, attention backward, per-launch GEMM timing,
# the ncu launch list of the bench command, ncu --set full of the named GEMM launches and the attention kernels, the
# compute-sanitizer passes, and - when tools/ubench/build_gemm_ab.sh has put binaries under gpurun_tmp/gemm_ab/ - the named
# GEMM launches through those historical kernel versions. Outputs: gpurun_out/${TAG}_*.  Usage: TAG=r03_final bash tools/final_validation.sh
TAG="${TAG:-r02_final}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
AFB_RECORD_ERRORS=gpurun_out/recorded_errors_final.json timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --mode train --steps 2 --warmup 3 > gpurun_out/${TAG}_train.json 2> gpurun_out/${TAG}_train.err
timeout 300 python bench.py --model qwen --steps 3 --warmup 3 --no-torch-cuda --no-cpu-full > gpurun_out/${TAG}_qwen.json 2> gpurun_out/${TAG}_qwen.err
timeout 120 python tools/vae_time.py 8 > gpurun_out/${TAG}_vae_b8.json 2>/dev/null
timeout 120 python tools/vae_time.py 1 > gpurun_out/${TAG}_vae_b1.json 2>/dev/null
timeout 120 python tools/diag_attn_bwd_time.py > gpurun_out/${TAG}_attn_bwd.json 2>/dev/null
timeout 200 python tools/profile_gemm_shapes.py --reps 5 --json gpurun_out/${TAG}_gemm_shapes.json > /dev/null 2> gpurun_out/${TAG}_shapes.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6500 --launch-count 2600 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-torch-cuda --no-cpu-full > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:gemm_bf16 -c 9 -o gpurun_out/${TAG}_gemm_full python tools/profile_gemm_shapes.py --reps 1 --no-warmup --once > gpurun_out/${TAG}_ncu_gemm.log 2>&1
ncu -i gpurun_out/${TAG}_gemm_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_gemm_shapes.csv 2>/dev/null; rm -f gpurun_out/${TAG}_gemm_full.ncu-rep
timeout 400 ncu --set full --clock-control none -k regex:attention --launch-skip 12 --launch-count 3 -o gpurun_out/${TAG}_attn_full python tools/diag_attn_bwd_time.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
ncu -i gpurun_out/${TAG}_attn_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_attention.csv 2>/dev/null; rm -f gpurun_out/${TAG}_attn_full.ncu-rep
SANITIZE_TIMEOUT=400 timeout 900 bash tools/sanitize.sh memcheck > gpurun_out/${TAG}_sanitize.txt 2>&1
SANITIZE_TIMEOUT=400 timeout 900 bash tools/sanitize.sh synccheck >> gpurun_out/${TAG}_sanitize.txt 2>&1
if ls gpurun_tmp/gemm_ab/gemm_ab_* > /dev/null 2>&1; then : > gpurun_out/${TAG}_gemm_ab.jsonl; for pass in 1 2; do for b in gpurun_tmp/gemm_ab/gemm_ab_*; do $b "$(basename $b | sed s/gemm_ab_//)" 5 >> gpurun_out/${TAG}_gemm_ab.jsonl; done; done; fi
cat gpurun_out/${TAG}_tests.log; cut -c1-160 gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_train.json gpurun_out/${TAG}_qwen.json; cat gpurun_out/${TAG}_sanitize.txt gpurun_out/${TAG}_attn_bwd.json
