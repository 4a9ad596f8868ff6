#!/bin/bash
# Standard GPU evidence pass (run under gpurun): parity tests, bench (both arms), ncu launch lists, ncu full captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'      (SKIP_TESTS=1 / SKIP_BENCH=1 / SKIP_NCU=1 to drop parts)
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -3 $O/pytest_$TAG.log
fi
if [ -z "$SKIP_BENCH" ]; then
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "bench ref rc=$?"
python bench.py --steps ${STEPS:-100} --warmup 5 > $O/bench_ours_$TAG.json 2> $O/bench_ours_$TAG.err; echo "bench ours rc=$?"
python - <<PY
import json
d = json.load(open('$O/bench_ours_$TAG.json')); r = json.load(open('$O/bench_ref_$TAG.json'))
print('C2 window: ours %.3f ms (e2e %.3f ms)  ref %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['ms_per_step']))
print(' roofline', d['roofline']['kernel'], d['roofline']['frac'], {k: v['ms'] for k, v in d['roofline']['stages_one_view'].items()})
c = d.get('also_C1')
if c: print('C1 step: %.4f ms (e2e %.4f ms)' % (c['ms_per_step'], c['e2e']['ms_per_step']), c['roofline']['frac'], {k: v['ms'] for k, v in c['roofline']['stages'].items()})
print(' cpu', d.get('cpu_baseline'))
PY
fi
if [ -z "$SKIP_NCU" ]; then
for WL in C1_tum_tracking C2_replica_mapping; do
S=${WL:0:2}
python tools/profile_step.py $WL 4 > $O/plain_${S}_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(preprocess|render|scatter|tile_sort|mark|home|bucket)' -s 5 -c 10 --csv --log-file $O/launches_${S}_$TAG.csv python tools/profile_step.py $WL 4 > $O/ncu_launch_${S}_$TAG.log 2>&1
echo "$S launch list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:^(preprocess|render|scatter|tile_sort|mark|home|bucket)' -s 12 -c 6 -f -o $O/prof_${S}_$TAG python tools/profile_step.py $WL 4 > $O/ncu_full_${S}_$TAG.log 2>&1
echo "$S full capture rc=$?"
done
fi
if [ -n "$WITH_SMOKE" ]; then python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log; fi
