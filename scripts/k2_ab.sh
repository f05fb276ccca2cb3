#!/bin/bash
# K2 A/B: bulk-copy mark kernel vs the register-load one, plus the contour test suite on the bulk path
python -m pytest tests/test_gpu_contour.py tests/test_abi_host.py -m gpu -x -q 2>&1 | tail -3
for res in 32768 16384; do
  echo "== bulk, res $res"; python scripts/k2_run.py --res $res --max_iter 10000 --reps 4 | tail -2
  echo "== register loads, res $res"; LM_K2_NO_BULK=1 python scripts/k2_run.py --res $res --max_iter 10000 --reps 4 | tail -2
done
