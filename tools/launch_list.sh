#!/bin/bash
# per-launch kernel times of one bench command (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: shares, not absolutes)
# usage: tools/launch_list.sh <out prefix under gpurun_out/> [bench.py args]
out=$1; shift
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-strong --no-cpu-baseline --no-wrapper "$@" > gpurun_out/${out}_plain.json 2> gpurun_out/${out}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${out}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-strong --no-cpu-baseline --no-wrapper "$@" > gpurun_out/${out}_ncu.log 2>&1
