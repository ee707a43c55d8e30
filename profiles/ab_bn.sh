#!/bin/bash
# A/B the N-tile width of the streaming (K-major) GEMMs via the GLF_DEBUG_BN tuning override
for bn in 128 256; do
  GLF_DEBUG_BN=$bn timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > /tmp/ab_$bn.json
  python -c "import json; d=json.load(open('/tmp/ab_$bn.json')); print('BN', $bn, d['value'], d['ms_per_step'])"
done
