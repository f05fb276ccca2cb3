#!/bin/bash
# Tuning sweep for K1 on the GPU box: rebuild with different -D tunables and time cfg3/cfg2.
set -u
cd "$(dirname "$0")/.."
for defs in "" "-DLM_K1_FB=64" "-DLM_K1_FB=48" "-DLM_K1_FB=64 -DLM_K1_COOL_MIN=8" "-DLM_K1_MIN_CTAS=2" "-DLM_K1_WARPS=4 -DLM_K1_MIN_CTAS=6"; do
  echo "=== defs: '$defs'"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build > /dev/null 2>&1 || echo BUILD FAILED
  LM_NVCC_DEFS="$defs" python scripts/k1_run.py --res 32768 --max_iter 10000 --reps 2 | tail -1
  LM_NVCC_DEFS="$defs" python scripts/k1_run.py --res 8192 --max_iter 2000 --reps 3 | tail -1
done
