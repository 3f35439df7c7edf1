#!/bin/bash
# GPU-box script (round 2): tests, bench, gather probe (times + dram bytes), C4 bench, ncu --set full of
# the parity-mode filter kernels.   gpurun --timeout 1500 -- 'bash tools/r2_probe.sh <tag>'
set -u
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 1500 $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
python bench.py --workload holter --steps 5 --warmup 3 > $O/${TAG}_bench_holter.json 2> $O/${TAG}_bench_holter.err
python tools/gather_probe.py > $O/${TAG}_gather_probe.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file $O/${TAG}_gather_probe_ncu.csv python tools/gather_probe.py > $O/${TAG}_gather_probe_ncu.log 2>&1
python tools/floor_only.py 3600 2 > $O/${TAG}_plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_sos_scan' -s 0 -c 4 \
    -f -o $O/${TAG}_full_sos python tools/floor_only.py 3600 2 > $O/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $O/${TAG}_full_sos.ncu-rep > $O/${TAG}_ncu_summary_sos.txt 2>&1
head -30 $O/${TAG}_gather_probe.txt
