#!/bin/bash
# ncu --set full capture (with SASS-level counters) of one launch of the packed extension kernel on the bench batch.
# usage: tools/ncu_extend.sh <out prefix under gpurun_out/> [bench.py args]
out=$1; shift
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-strong --no-cpu-baseline --no-wrapper "$@" > gpurun_out/${out}_plain.json 2> gpurun_out/${out}_plain.err || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_extend<\(bool\)0, \(bool\)1>' -s 1 -c 1 -f \
    -o gpurun_out/${out} python bench.py --steps 1 --warmup 1 --no-strong --no-cpu-baseline --no-wrapper "$@" > gpurun_out/${out}_ncu.log 2>&1
ls -la gpurun_out/${out}.ncu-rep
