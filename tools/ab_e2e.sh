#!/bin/bash
# same-box A/B of the end-to-end (host buffers) number:  tools/ab_e2e.sh "NAME:ENV=V" ...
for rep in 1 2; do
  for spec in "$@"; do
    name="${spec%%:*}"; envs="${spec#*:}"
    env $envs timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /tmp/ab_out.txt 2>&1
    python -c "
import json,sys
for l in open('/tmp/ab_out.txt'):
    if l.startswith('{'):
        d=json.loads(l); print('%-10s value %.1f (%.2f ms)  e2e %.1f (%.2f ms)  clk %s' % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])); break
else: print(sys.argv[1],'NO JSON')
" "$name"
  done
done
