#!/bin/bash
# Evidence for the int8 tier of the headline step: plain run, ncu launch list, one --set full capture of the scan kernels.
set -x
mkdir -p gpurun_out
python bench.py --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r2_int8_plain.json 2> gpurun_out/r2_int8_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_int8_launches.csv \
    python bench.py --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r2_int8_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 40 -c 4 -o gpurun_out/r2_int8_full -f \
    python bench.py --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r2_int8_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
