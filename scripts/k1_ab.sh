#!/bin/bash
# K1 A/B of the escape handling (LM_K1_ESC_MODE) and the adaptive cool-down on configs 2, 3, 1 (dwell only) and 2 with the potential
for defs in "-DLM_K1_ESC_MODE=1 -DLM_K1_ADAPTIVE_COOL=0" "-DLM_K1_ESC_MODE=2 -DLM_K1_ADAPTIVE_COOL=0" "-DLM_K1_ESC_MODE=1 -DLM_K1_ADAPTIVE_COOL=1" "-DLM_K1_ESC_MODE=2 -DLM_K1_ADAPTIVE_COOL=1" "$@"; do
  echo "== $defs"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null || { echo build failed; continue; }
  LM_NVCC_DEFS="$defs" python scripts/k1_run.py --res 8192 --max_iter 2000 --reps 4 | tail -2
  LM_NVCC_DEFS="$defs" python scripts/k1_run.py --res 8192 --max_iter 2000 --reps 3 --field 1 | tail -1
  LM_NVCC_DEFS="$defs" python scripts/k1_run.py --res 2000 --max_iter 500 --reps 4 | tail -1
  LM_NVCC_DEFS="$defs" python scripts/k1_run.py --res 32768 --max_iter 10000 --reps 2 | tail -1
done
python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null
python -m pytest tests/test_gpu_escape.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
