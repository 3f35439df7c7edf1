#!/bin/bash
# Runs ON the GPU box (via gpurun): tests, both bench arms, per-launch ncu list, ncu --set full of
# the top kernels.  Everything lands in gpurun_out/ev_*; tools/collect_profiles.py (run on the
# authoring box) copies the summaries into profiles/.
#   gpurun --timeout 1500 -- 'bash tools/capture_evidence.sh <tag>'
set -u
TAG=${1:-v0}
O=gpurun_out
mkdir -p $O
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/ev_${TAG}_pytest_gpu.log
  cat $O/ev_${TAG}_pytest_gpu.log
fi
python bench.py --impl reference --steps 3 --warmup 1 > $O/ev_${TAG}_bench_reference.json 2> $O/ev_${TAG}_bench_reference.err
python bench.py --dump-kernels $O/ev_${TAG}_kernels_parity.json > $O/ev_${TAG}_bench_parity.json 2> $O/ev_${TAG}_bench_parity.err
python bench.py --filter-mode fullrate --no-cpu-baseline --dump-kernels $O/ev_${TAG}_kernels_fullrate.json > $O/ev_${TAG}_bench_fullrate.json 2> $O/ev_${TAG}_bench_fullrate.err
python bench.py --e2e-ingest sm --no-cpu-baseline --no-profile > $O/ev_${TAG}_bench_parity_ingest_sm.json 2> $O/ev_${TAG}_bench_parity_ingest_sm.err
python bench.py --workload holter --steps 5 --warmup 3 > $O/ev_${TAG}_bench_holter_c4.json 2> $O/ev_${TAG}_bench_holter_c4.err
python bench.py --workload batch --steps 5 --warmup 3 > $O/ev_${TAG}_bench_batch_c3.json 2> $O/ev_${TAG}_bench_batch_c3.err
timeout 120 python tools/ingest_probe.py > $O/ev_${TAG}_ingest_probe.json 2> $O/ev_${TAG}_ingest_probe.err
cut -c1-300 $O/ev_${TAG}_bench_parity.json
# per-launch device times of the bench command (cold-cache, serialised: compare SHARES)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > $O/ev_${TAG}_plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ev_${TAG}_launches_parity.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > $O/ev_${TAG}_ncu_launches.log 2>&1
# top kernels, full sections (short driver: two stage-A calls on the C2 recording)
python tools/floor_only.py 3600 2 > $O/ev_${TAG}_plain_full_parity.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_rolling_floor_blk|k_scan' -s 0 -c 4 \
    -f -o $O/ev_${TAG}_full_parity python tools/floor_only.py 3600 2 > $O/ev_${TAG}_ncu_full_parity.log 2>&1
python tools/floor_only.py 3600 2 fullrate > $O/ev_${TAG}_plain_full_fullrate.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_contract_i16|k_scan' -s 0 -c 3 \
    -f -o $O/ev_${TAG}_full_fullrate python tools/floor_only.py 3600 2 fullrate > $O/ev_${TAG}_ncu_full_fullrate.log 2>&1
python tools/ncu_summary.py $O/ev_${TAG}_full_parity.ncu-rep > $O/ev_${TAG}_ncu_summary_parity.txt 2>&1
python tools/ncu_summary.py $O/ev_${TAG}_full_fullrate.ncu-rep > $O/ev_${TAG}_ncu_summary_fullrate.txt 2>&1
tail -3 $O/ev_${TAG}_ncu_full_parity.log
du -sh $O; ls -la $O | grep ev_${TAG}
