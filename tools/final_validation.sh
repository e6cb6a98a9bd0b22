mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
AFB_RECORD_ERRORS=gpurun_out/recorded_errors_final.json timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_c28_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c28_bench.json 2> gpurun_out/r02_c28_bench.err
timeout 300 python bench.py --mode train --steps 2 --warmup 3 > gpurun_out/r02_c28_train.json 2> gpurun_out/r02_c28_train.err
timeout 300 python bench.py --model qwen --steps 3 --warmup 3 --no-torch-cuda --no-cpu-full > gpurun_out/r02_c28_qwen.json 2> gpurun_out/r02_c28_qwen.err
timeout 120 python tools/vae_time.py 8 > gpurun_out/r02_c28_vae_b8.json 2>/dev/null
timeout 120 python tools/vae_time.py 1 > gpurun_out/r02_c28_vae_b1.json 2>/dev/null
timeout 120 python tools/diag_attn_bwd_time.py > gpurun_out/r02_c28_attn_bwd.json 2>/dev/null
timeout 200 python tools/profile_gemm_shapes.py --reps 5 --json gpurun_out/r02_c28_gemm_shapes.json > /dev/null 2> gpurun_out/r02_c28_shapes.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6500 --launch-count 2600 --csv --log-file gpurun_out/r02_c28_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-torch-cuda --no-cpu-full > gpurun_out/r02_c28_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:gemm_bf16 -c 9 -o gpurun_out/r02_c28_gemm_full python tools/profile_gemm_shapes.py --reps 1 --no-warmup --once > gpurun_out/r02_c28_ncu_gemm.log 2>&1
ncu -i gpurun_out/r02_c28_gemm_full.ncu-rep --page raw --csv > gpurun_out/r02_c28_ncu_full_gemm_shapes.csv 2>/dev/null; rm -f gpurun_out/r02_c28_gemm_full.ncu-rep
timeout 400 ncu --set full --clock-control none -k regex:attention --launch-skip 12 --launch-count 3 -o gpurun_out/r02_c28_attn_full python tools/diag_attn_bwd_time.py > gpurun_out/r02_c28_ncu_attn.log 2>&1
ncu -i gpurun_out/r02_c28_attn_full.ncu-rep --page raw --csv > gpurun_out/r02_c28_ncu_full_attention.csv 2>/dev/null; rm -f gpurun_out/r02_c28_attn_full.ncu-rep
SANITIZE_TIMEOUT=400 timeout 900 bash tools/sanitize.sh memcheck > gpurun_out/r02_c28_sanitize.txt 2>&1
SANITIZE_TIMEOUT=400 timeout 900 bash tools/sanitize.sh synccheck >> gpurun_out/r02_c28_sanitize.txt 2>&1
cat gpurun_out/r02_c28_tests.log; cut -c1-160 gpurun_out/r02_c28_bench.json gpurun_out/r02_c28_train.json gpurun_out/r02_c28_qwen.json; cat gpurun_out/r02_c28_sanitize.txt gpurun_out/r02_c28_attn_bwd.json
