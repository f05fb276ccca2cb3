#!/bin/bash
# A/B of the K3 build switches on the config-5 generator (4e6 polynomials)
for defs in "-DLM_K3_FP32_PHASE=0 -DLM_K3_SUM_RCP_STEPS=1" "-DLM_K3_FP32_PHASE=0 -DLM_K3_SUM_RCP_STEPS=0" "-DLM_K3_FP32_PHASE=1 -DLM_K3_SUM_RCP_STEPS=0" "$@"; do
  echo "== $defs"
  LM_NVCC_DEFS="$defs" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null
  LM_NVCC_DEFS="$defs" python scripts/k3_run.py 4000000 2>&1 | grep -v Warning | grep -v "lam =" | tail -4
done
python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null
