#!/bin/bash
# quick bench over alternative builds of the library: args = lib paths ("" = default); env QB_ARGS = bench args
for lib in "$@"; do
  echo "== lib=$lib $QB_ARGS"
  MSDA_B200_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra $QB_ARGS 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline_fwd_bwd']; print('%.1f Mq/s fwd %.4f ms bwd %.4f ms frac %.4f' % (d['value']/1e6, r['fwd_ms_per_layer'], r['bwd_ms_per_layer'], r['frac']))"
done
