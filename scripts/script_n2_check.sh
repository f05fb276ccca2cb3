#!/bin/bash
# the drop-in script on 1 GPU and, through torchrun, on 2 GPUs: the three output files must be identical
set -e
M=inverse_eigenvalue_loci_mandelbrot_correspondence_b200.mandelbrot_boundary_sample
ARGS="--xlim -2.1 0.9 --ylim -1.5 1.5 --res 4096 --max_iter 1000 --level 0.96"
python -m $M $ARGS --output_prefix /tmp/one/mandel
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 -m $M $ARGS --output_prefix /tmp/two/mandel
cmp /tmp/one/mandel_boundary.csv /tmp/two/mandel_boundary.csv && cmp /tmp/one/mandel_meta.txt /tmp/two/mandel_meta.txt && echo "CSV and meta identical: $(wc -l < /tmp/one/mandel_boundary.csv) lines"
