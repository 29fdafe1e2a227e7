#!/bin/bash
# Sweep of the tile kernel's scheduling knobs (chunk size, chunk order) on config 2.
for order in ${ORDERS:-1}; do for chunk in ${CHUNKS:-1 2 4 8}; do
    SNPG_TILE_ORDER=$order SNPG_TILE_CHUNK=$chunk timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('order $order chunk $chunk fused_us', round(d['kernels']['fused']['ms_per_step']*1000,1), 'step_ms', round(d['ms_per_step'],4))"
done; done
