#!/bin/bash
# K2 mark-kernel sweep: rebuild with different -D and time the three K2 kernels on a resident 32768^2 grid
for defs in "-DLM_K2_MARK_ROWS=8" "-DLM_K2_MARK_ROWS=6" "-DLM_K2_MARK_ROWS=3" ""; do
  echo "=== defs: '$defs'"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null || { echo build failed; continue; }
  LM_NVCC_DEFS="$defs" python scripts/k2_run.py --res 32768 --max_iter 10000 --reps 4 | tail -2
done
