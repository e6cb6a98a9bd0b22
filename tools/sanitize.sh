#!/usr/bin/env bash
# compute-sanitizer over the small-shape operator tests (the build must add memcheck / racecheck / synccheck
# targets). Runs on a GPU box:
#     tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck|all] [extra pytest args]
# Each tool runs the op-level parity tests at their smallest shapes (the kernels are the product kernels; the sanitizer
# slows launches 10-100x, so the full-size tests are left out). Output: gpurun_out/sanitize_<tool>.log, one summary line per
# tool on stdout, exit code 1 if any tool reports errors.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tool="${1:-all}"
shift || true
SAN="${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}"
[ -x "$SAN" ] || SAN=compute-sanitizer
# small shapes only: one GEMM per epilogue, attention with a ragged tail, the streaming kernels, the sampler, the backward ops
SELECT='test_gemm_parity or test_gemm_strided or test_attention_parity and not 4608 or test_attention_bounded and not 4608 or test_ln_modulate_parity or test_rmsnorm_rope_parity or test_small_linear_parity or test_sampler_step_matches or test_timestep_embed'
FILES="tests/test_gpu_parity.py"
BWD_SELECT='test_attention_backward_parity and not 1000 or test_gemm_transposed_weight or test_rowscale_and_gelu or test_ln_modulate_bwd or test_lora_dropout_mask'
OPT_SELECT='adamw8bit or skips'
VAE_SELECT='test_conv3x3_implicit_gemm and not 64-48 or test_groupnorm_swish and not 64-64 or test_upsample_softmax'
tools=("$tool")
[ "$tool" = all ] && tools=(memcheck racecheck synccheck)
rc=0
for t in "${tools[@]}"; do
  log="gpurun_out/sanitize_${t}.log"
  : > "$log"
  for spec in "tests/test_gpu_parity.py|$SELECT" "tests/test_gpu_backward.py|$BWD_SELECT" "tests/test_gpu_vae.py|$VAE_SELECT" \
              "tests/test_gpu_optim.py|$OPT_SELECT"; do
    f="${spec%%|*}"; k="${spec#*|}"
    echo "=== $t: $f -k '$k'" >> "$log"
    timeout "${SANITIZE_TIMEOUT:-900}" "$SAN" --tool "$t" --error-exitcode 66 --print-limit 20 --launch-timeout 120 \
      python -m pytest "$f" -q -x -m gpu -k "$k" -p no:cacheprovider "$@" >> "$log" 2>&1
    code=$?
    echo "=== exit $code" >> "$log"
    [ $code -ne 0 ] && rc=1
  done
  errs=$(grep -c "========= .*\(Invalid\|Race\|hazard\|Barrier error\|Uninitialized\)" "$log" || true)
  summ=$(grep "ERROR SUMMARY" "$log" | tr '\n' ';')
  echo "sanitize $t: reported_lines=$errs rc=$rc  $summ"
done
exit $rc
