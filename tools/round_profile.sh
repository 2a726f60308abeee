#!/bin/bash
# Diagnostic helper (not part of the product): the end-of-round measurement set on one B200 -
# GPU test suite, bench lines of every workload, reference arm, ncu launch list and one --set full capture per dominant kernel.
# usage (on the GPU box, from the repo root): bash tools/round_profile.sh TAG
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${TAG}_gputests.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_c4_reference.json 2> $O/${TAG}_bench_c4_reference.err
python bench.py --workload C5 --steps 5 --warmup 3 --sweep > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
for w in C3 C2 C1; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > $O/${TAG}_bench_${w}.json 2> $O/${TAG}_bench_${w}.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench_c4.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_q -s 5 -c 1 -f -o $O/${TAG}_c4_step \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgemm -s 2 -c 1 -f -o $O/${TAG}_c5_bgemm \
    python bench.py --workload C5 --total-batch 512 --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_c5.log 2>&1
python tools/small_probe.py > $O/${TAG}_small_probe.log 2>&1
python tools/kernel_zoo.py > $O/${TAG}_zoo_timings.jsonl 2> $O/${TAG}_zoo.err
ncu --set full --clock-control none --import-source on -k regex:step_small -s 30 -c 1 -f -o $O/${TAG}_c1_small \
    python tools/small_probe.py > $O/${TAG}_ncu_c1.log 2>&1
cat $O/${TAG}_small_probe.log
cat $O/${TAG}_gputests.log
tail -c 400 $O/${TAG}_bench_c4.json
ls -la $O/${TAG}_*
