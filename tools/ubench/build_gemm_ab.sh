#!/usr/bin/env bash
# Builds tools/ubench/gemm_ab.cu against historical versions of csrc/gemm.cu (one binary per git revision) into
# gpurun_tmp/gemm_ab/ (git-ignored, travels to the GPU box). Usage: tools/ubench/build_gemm_ab.sh tag=rev [tag=rev ...]
set -eu
cd "$(dirname "$0")/../.."
out=gpurun_tmp/gemm_ab
mkdir -p "$out"
for spec in "$@"; do
  tag="${spec%%=*}"; rev="${spec#*=}"
  src=/tmp/gemm_ab_src/$tag
  rm -rf "$src"; mkdir -p "$src/include" "$src/arcflow_b200/csrc"
  git show "$rev:include/arcflow_b200.h" > "$src/include/arcflow_b200.h"
  for f in gemm.cu host.cu common.cuh; do git show "$rev:arcflow_b200/csrc/$f" > "$src/arcflow_b200/csrc/$f"; done
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I"$src" -o "$out/gemm_ab_$tag" \
    tools/ubench/gemm_ab.cu "$src/arcflow_b200/csrc/gemm.cu" "$src/arcflow_b200/csrc/host.cu" -lcuda 2>&1 | grep -v deprecat || true
  ls -la "$out/gemm_ab_$tag" | cut -c25-
done
