#!/bin/bash
cd "$(dirname "$0")/.."
for f in 0 8; do
  RTX_FLAT_ITEMS=$f python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('flat_items<=$f', 'Mrays/s %.0f ms %.1f share %s' % (d['value'], d['ms_per_step'], {k: round(v,3) for k,v in r['kernel_share_of_step'].items()}))"
done
