#!/bin/bash
# Standard GPU evidence pass (run under gpurun): parity tests, bench, ncu launch list, ncu full capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
if [ -z "$ONLY_NCU" ]; then
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -3 $O/pytest_$TAG.log
python bench.py --impl reference --steps 30 --warmup 5 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "bench ref rc=$?"
python bench.py --steps 200 --warmup 5 > $O/bench_ours_$TAG.json 2> $O/bench_ours_$TAG.err; echo "bench ours rc=$?"
cat $O/bench_ours_$TAG.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stages'], d.get('cpu_baseline'))"
fi
python tools/profile_step.py C1_tum_tracking 3 > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(preprocess|render|scatter|tile_sort|mark)' -s 5 -c 8 --csv --log-file $O/launches_$TAG.csv python tools/profile_step.py C1_tum_tracking 3 > $O/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/profile_step.py C1_tum_tracking 3 > $O/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:^(preprocess|render|scatter|tile_sort|mark)' -s 9 -c 4 -f -o $O/prof_$TAG python tools/profile_step.py C1_tum_tracking 3 > $O/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
