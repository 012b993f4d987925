#!/bin/bash
# usage: tools/ab_env.sh VAR v1 v2 ...   -- bench.py value / ms per step for each setting of one environment switch
var=$1; shift
for v in "$@"; do
  env "$var=$v" python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$var=$v', round(d['value'],2), round(d['ms_per_step'],3), d['aux'])"
done
