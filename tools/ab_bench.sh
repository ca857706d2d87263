#!/bin/bash
# run bench.py (device-resident leg only matters) for every variant library given; prints ms/step, chain ms, level-0 ms, crc
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset SHINESTACK_B200_LIB; else export SHINESTACK_B200_LIB=$PWD/shinestacker_b200/variants/lib_$lib.so; fi
  python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-next-rows 2>&1 | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
try:
    d=json.loads(t[-1]); print('$lib', round(d['ms_per_step'],3), round(d['device_chain_ms_per_step'],3), round(d['roofline']['launch_ms'],3), d['output_crc32'])
except Exception as e:
    print('$lib', 'FAILED', t[-3:])
"
done
