#!/bin/bash
# sweep of the DMA-ring tunables for the fused Atari-shaped sample (bench.py --only c3): step time by events
mkdir -p gpurun_out
for cfg in "4 4 8192" "6 4 8192" "8 2 8192" "8 3 8192" "5 4 8192" "4 8 4096" "8 4 4096" "6 6 4096" "4 4 16384" "3 4 8192"; do
  set -- $cfg
  MRL_BULK_CTAS_PER_SM=$1 MRL_BULK_STAGES=$2 MRL_BULK_CHUNK=$3 timeout 200 python bench.py --only c3 --steps 100 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ctas/sm $1 stages $2 chunk $3: step us %.2f' % (d['ms_per_step']*1e3), 'parity', d['parity_check']['ok'])
"
done
