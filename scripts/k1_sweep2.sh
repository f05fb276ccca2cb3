#!/bin/bash
# K1 tunables on configs 2 (dwell only / with potential), 1 and 3
for defs in "" "-DLM_K1_FB=32" "-DLM_K1_FB=48" "-DLM_K1_FB=96" "-DLM_K1_COOL_MIN=2" "-DLM_K1_COOL_MIN=8" "-DLM_K1_MIN_CTAS=4" "-DLM_K1_WARPS=4 -DLM_K1_MIN_CTAS=6"; do
  echo "== '$defs'"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null || { echo build failed; continue; }
  LM_NVCC_DEFS="$defs" python scripts/k1_dev_run.py --res 8192 --max_iter 2000 --reps 4 | tail -1
  LM_NVCC_DEFS="$defs" python scripts/k1_dev_run.py --res 2000 --max_iter 500 --reps 6 | tail -1
  LM_NVCC_DEFS="$defs" python scripts/k1_dev_run.py --res 32768 --max_iter 10000 --reps 2 | tail -1
done
python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null
