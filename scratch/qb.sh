#!/bin/bash
# quick bench: prints Mq/s, fwd ms, bwd ms, frac for each arg string
for a in "$@"; do
  echo "== bench $a"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra $a 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline_fwd_bwd']; print('%.1f Mq/s fwd %.4f ms bwd %.4f ms frac %.4f' % (d['value']/1e6, r['fwd_ms_per_layer'], r['bwd_ms_per_layer'], r['frac']))"
done
