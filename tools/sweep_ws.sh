#!/bin/bash
for w in "$@"; do
  GNB_WS_GIB=$w python bench.py --steps 4 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ws_gib', '$w', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'grint', round(d['secondary']['value']), 'n512', round(d['secondary_n512']['value']))"
done
