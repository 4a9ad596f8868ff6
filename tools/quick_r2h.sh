#!/bin/bash
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-also-c1 --steps 40 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['details']['units_per_rank'][0])
PY
}
GSR_STREAM_PRIO=1 $B --as-rank-of 8 > $O/r2h_p8.json 2>$O/r2h.err; show $O/r2h_p8.json "rank0-of-8 prio"
$B --as-rank-of 8 > $O/r2h_8.json 2>$O/r2h.err; show $O/r2h_8.json "rank0-of-8"
GSR_STREAM_PRIO=1 $B --as-rank-of 4 > $O/r2h_p4.json 2>$O/r2h.err; show $O/r2h_p4.json "rank0-of-4 prio"
GSR_STREAM_PRIO=1 $B > $O/r2h_p1.json 2>$O/r2h.err; show $O/r2h_p1.json "N=1 prio"
$B --engines 5 > $O/r2h_e5.json 2>$O/r2h.err; show $O/r2h_e5.json "N=1 engines 5"
$B --engines 3 > $O/r2h_e3.json 2>$O/r2h.err; show $O/r2h_e3.json "N=1 engines 3"
tail -3 $O/r2h.err
