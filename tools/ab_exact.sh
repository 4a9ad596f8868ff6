#!/usr/bin/env bash
# A/B of the exact-exp variants on the GPU box: gradient error vs the reference and C1 / C2 step time.
set -u
cd "$(dirname "$0")/.."
for div in 0 1; do
  GSR_EXTRA_NVCC_FLAGS="-DGSR_BWD_DIV=$div" python gs-slam-analytica_jacobian_b200/build.py --force > /dev/null
  echo "=== GSR_BWD_DIV=$div (exact backward: expf + $([ $div = 1 ] && echo 'correctly rounded division' || echo 'rcp.approx'))"
  python tools/grad_noise.py C3_batched_tracking 2>&1 | tail -8
  for mode in 2 1 -1; do
    echo "--- GSR_EXACT_EXP=$mode C1"
    GSR_EXACT_EXP=$mode python bench.py --workload C1_tum_tracking --no-cpu-baseline --steps 100 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=d['roofline']['stages']
print('ms/step %.4f e2e %.4f | fwd %.4f bwd %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], st['render_forward']['ms'], st['render_backward']['ms']))"
  done
  [ $div = 1 ] && break
done
GSR_EXTRA_NVCC_FLAGS="" python gs-slam-analytica_jacobian_b200/build.py --force > /dev/null
for mode in 2 1 -1; do
  echo "--- GSR_EXACT_EXP=$mode C2 (DIV=0)"
  GSR_EXACT_EXP=$mode python bench.py --no-cpu-baseline --no-also-c1 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=d['roofline']['stages_one_view']
print('ms/step %.4f e2e %.4f | fwd %.4f bwd %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], st['render_forward']['ms'], st['render_backward']['ms']))"
done
