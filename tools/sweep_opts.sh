#!/bin/bash
# dev helper: bench value for a list of GNB_DEV_OPTS settings:  bash tools/sweep_opts.sh "rk_sms=132" "rk_sms=120,rec_streams=3"
for o in "$@"; do
  GNB_DEV_OPTS="$o" python bench.py --steps 4 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$o', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'grint', round(d['secondary']['value']), 'n512', round(d['secondary_n512']['value']))"
done
