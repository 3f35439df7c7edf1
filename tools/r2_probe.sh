#!/bin/bash
# GPU-box script: gather probe (times + dram bytes) and ncu --set full of the parity-mode filter kernels.
set -u
O=gpurun_out
mkdir -p $O
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_probe.cu -o /tmp/gather_probe || exit 1
/tmp/gather_probe > $O/r2b_gather_probe.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file $O/r2b_gather_probe_ncu.csv /tmp/gather_probe > $O/r2b_gather_probe_ncu.log 2>&1
python bench.py --workload holter --steps 5 --warmup 3 > $O/r2b_bench_holter.json 2> $O/r2b_bench_holter.err
python tools/floor_only.py 3600 2 > $O/r2b_plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_sos_scan' -s 0 -c 4 \
    -f -o $O/r2b_full_sos python tools/floor_only.py 3600 2 > $O/r2b_ncu_full.log 2>&1
python tools/ncu_summary.py $O/r2b_full_sos.ncu-rep > $O/r2b_ncu_summary_sos.txt 2>&1
cat $O/r2b_gather_probe.txt | head -40
