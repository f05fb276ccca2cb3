#!/bin/bash
for defs in "-DLM_K4A_CPT=8" "-DLM_K4A_CPT=6" "-DLM_K4A_CPT=2" ""; do
  echo "=== defs: '$defs'"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null || { echo build failed; continue; }
  LM_NVCC_DEFS="$defs" python scripts/k4a_run.py 1000000 | grep "variant 0" | tail -1
done
