#!/bin/bash
# ncu sweep of the K2 schedule knobs on the 8-GPU shard shape (1.1M docs x 4096, 10k queries): DRAM bytes, L2 hit rate,
# tensor-pipe activity and duration per configuration.  Usage (under gpurun): bash tools/l2_sweep.sh "<env assignments>" ...
export LR_B200_NO_BUILD=1
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,sm__cycles_elapsed.avg.per_second"
for cfg in "$@"; do
  echo "=== $cfg"
  export $cfg
  ncu --metrics $M --clock-control none -k regex:umma_gemm -s 3 -c 1 --csv python bench.py --docs ${DOCS:-1100000} --queries ${QUERIES:-10000} --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
for r in rows[1:]:
    print('   ', r[-3], r[-2], r[-1])
"
  for kv in $cfg; do unset ${kv%%=*}; done
done
