#!/bin/bash
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-also-c1 --steps 40 --warmup 3"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['details']['units_per_rank'][0])
PY
}
for WB in 1 2 3; do for E in 2 3 4; do
$B --as-rank-of 8 --whole-bands $WB --engines $E > $O/r2g_$WB$E.json 2>$O/r2g.err; show $O/r2g_$WB$E.json "rank0-of-8 wb=$WB engines=$E"
done; done
$B --as-rank-of 4 --whole-bands 1 > $O/r2g_4.json 2>$O/r2g.err; show $O/r2g_4.json "rank0-of-4 wb=1"
$B --as-rank-of 4 --whole-bands 2 > $O/r2g_42.json 2>$O/r2g.err; show $O/r2g_42.json "rank0-of-4 wb=2"
tail -3 $O/r2g.err
