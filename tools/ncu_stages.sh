#!/bin/bash
# ncu --set full of one launch each of k_seeds, k_cluster and the packed k_extend on a bench batch (the step after the warm-up).
# usage: tools/ncu_stages.sh <out prefix under gpurun_out/> [bench.py args]
out=$1; shift
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-strong --no-cpu-baseline --no-wrapper "$@" > gpurun_out/${out}_plain.json 2> gpurun_out/${out}_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:k_seeds|k_cluster|k_extend' -s 4 -c 3 -f -o gpurun_out/${out} \
    python bench.py --steps 1 --warmup 1 --no-strong --no-cpu-baseline --no-wrapper "$@" > gpurun_out/${out}_ncu.log 2>&1
ls -la gpurun_out/${out}.ncu-rep
