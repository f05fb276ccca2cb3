#!/bin/bash
# K1 tunables sweep (rebuilds the library with -D overrides; see csrc/lm_escape.cu)
for defs in "-DLM_K1_FB=64 -DLM_K1_COOL_MIN=4" "-DLM_K1_FB=96 -DLM_K1_COOL_MIN=4" "-DLM_K1_FB=128 -DLM_K1_COOL_MIN=4" "-DLM_K1_FB=128 -DLM_K1_COOL_MIN=8" "-DLM_K1_FB=64 -DLM_K1_COOL_MIN=16" "-DLM_K1_FB=64 -DLM_K1_COOL_MIN=8"; do
  echo "=== defs: '$defs'"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null || { echo build failed; continue; }
  LM_NVCC_DEFS="$defs" python scripts/k1_run.py --res 8192 --max_iter 2000 --reps 3 | tail -1
  LM_NVCC_DEFS="$defs" python scripts/k2_run.py --res 32768 --max_iter 10000 --reps 1 > /dev/null
  LM_NVCC_DEFS="$defs" python bench.py --no-roots --no-cpu-baseline --no-e2e --steps 3 2>&1 | tail -1 | cut -c1-110
done
