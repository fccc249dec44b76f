#!/bin/bash
# same-box A/B of bench.py under different env settings:  tools/ab.sh "NAME1:ENV1=V1" "NAME2:ENV2=V2" ...
show() { python -c "
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); f=d['roofline']['family_ms_per_step']
        print('%-10s %.1f pairs/s %.2f ms  lin %.2f conv %.2f attn %.2f ln %.2f  clk %s' % (sys.argv[2], d['value'], d['ms_per_step'], f['gemm_linear'], f['gemm_conv3x3'], f['attention'], f['layernorm'], d['clocks']['sm_mhz']))
        break
else:
    print(sys.argv[2], 'NO JSON'); print(open(sys.argv[1]).read()[-600:])
" "$1" "$2"; }
for rep in 1 2; do
  for spec in "$@"; do
    name="${spec%%:*}"; envs="${spec#*:}"
    env $envs timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /tmp/ab_out.txt 2>&1
    show /tmp/ab_out.txt "$name"
  done
done
