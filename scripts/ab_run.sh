# developer A/B: device-resident refine of a few workload shapes under different env settings
# usage: SHAPES="map50:60 c5a:12" bash scripts/ab_run.sh [ENV=VAL ...]   (one run per argument, plus the default)
run() { tag=$1; shift; for w in $SHAPES; do set -- ${w/:/ }; env "${ENVV[@]}" timeout 300 python scripts/dev_shape.py $1 $2 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', d['workload'], '%.0f QP/s'%d['qp_per_s'], d['launch'])"; done; }
SHAPES=${SHAPES:-"c5c:12 map50:60 c5:16 c5a:12"}
ENVV=(A=1); run default
for e in "$@"; do ENVV=($e); run $e; done
