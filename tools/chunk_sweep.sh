#!/bin/bash
cd "$(dirname "$0")/.."
for c in 1048576 2097152 4194304 8388608; do
  RTX_CHUNK=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('chunk $c', 'Mrays/s %.0f ms %.1f e2e %.0f share %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v,3) for k,v in r['kernel_share_of_step'].items()}))"
done
