#!/bin/bash
# K2 mark kernel: time and DRAM traffic per launch for different row-group heights (ncu, one pass of metrics)
for rows in 4 6 7 8; do
  echo "=== LM_K2_MARK_ROWS=$rows"
  LM_NVCC_DEFS="-DLM_K2_MARK_ROWS=$rows" python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null || { echo build failed; continue; }
  LM_NVCC_DEFS="-DLM_K2_MARK_ROWS=$rows" python scripts/k2_run.py --res 32768 --max_iter 10000 --reps 3 | tail -1
  LM_NVCC_DEFS="-DLM_K2_MARK_ROWS=$rows" ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:contour_mark -c 3 --csv python scripts/k2_run.py --res 32768 --max_iter 10000 --reps 3 2>/dev/null | grep -E "contour_mark" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | tail -8
done
python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build --force > /dev/null
