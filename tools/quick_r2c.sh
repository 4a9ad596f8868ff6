#!/bin/bash
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-also-c1 --steps 40 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:v['ms'] for k,v in d['roofline']['stages_one_view'].items()})
PY
}
$B --as-rank-of 8 > $O/r2c_rank8.json 2>$O/r2c.err; show $O/r2c_rank8.json "C2 rank0-of-8 share"
$B --as-rank-of 8 --engines 1 > $O/r2c_rank8_e1.json 2>>$O/r2c.err; show $O/r2c_rank8_e1.json "C2 rank0-of-8 share, 1 engine"
GSR_SCATTER_DIRECT=1 $B > $O/r2c_direct.json 2>>$O/r2c.err; show $O/r2c_direct.json "C2 direct-atomic scatter"
$B > $O/r2c_base.json 2>>$O/r2c.err; show $O/r2c_base.json "C2 base"
GSR_SCATTER_DIRECT=1 $B --workload C4_large --views 2 --steps 10 > $O/r2c_c4_direct.json 2>>$O/r2c.err; show $O/r2c_c4_direct.json "C4 v2 direct-atomic scatter"
$B --workload C4_large --views 2 --steps 10 > $O/r2c_c4.json 2>>$O/r2c.err; show $O/r2c_c4.json "C4 v2 base"
tail -3 $O/r2c.err
