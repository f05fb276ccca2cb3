#!/bin/bash
for defs in "-DLM_K4_PF_SMOOTH=1 -DLM_K4_PF_LAP=2" "-DLM_K4_PF_SMOOTH=2 -DLM_K4_PF_LAP=4" "-DLM_K4_PF_SMOOTH=4 -DLM_K4_PF_LAP=8" "-DLM_K4_PF_SMOOTH=8 -DLM_K4_PF_LAP=6"; do
  echo "=== defs: '$defs'"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null || { echo build failed; continue; }
  LM_NVCC_DEFS="$defs" python scripts/k4_time.py
done
