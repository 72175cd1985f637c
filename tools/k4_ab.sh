#!/bin/bash
# K4 A/B on one box: parity tests of the sparse path, then c4 / c4uniform with the round-1 kernel and the row kernel
mkdir -p gpurun_out/k4
timeout 600 python -m pytest tests -m gpu -q -x -k "sparse or hybrid or impact" 2>&1 | tail -4
IFS=","; for cfg in ${K4_CFGS:-3 16,3 32}; do IFS=" "
  kern=$(echo $cfg | cut -d' ' -f1); kb=$(echo $cfg | cut -d' ' -f2)
  for c in "c4uniform --docs 1100000" "c4 --docs 1100000 --steps 3" "c4uniform --docs 1100000 --queries 32" "c4 --docs 1100000 --queries 32"; do
    n=$(echo $c | tr -d ' -' )
    LR_SPARSE_KERNEL=$kern LR_SPARSE_STEP_KB=$kb timeout 600 python bench.py --config $c --warmup 3 --no-cpu-baseline > gpurun_out/k4/${n}_k${kern}_kb${kb}.json 2> gpurun_out/k4/${n}_k${kern}_kb${kb}.err
    echo "$n kernel=$kern kb=$kb rc=$? $(python -c "
import json,sys
d=json.load(open('gpurun_out/k4/${n}_k${kern}_kb${kb}.json'))
print(round(d['ms_per_step'],3),'ms frac',round(d['roofline']['frac'],4),'kernel_ms',round(d['roofline']['kernel_ms'],3),'parity',d['parity']['ok'])
" 2>&1 | tail -1) $(tail -1 gpurun_out/k4/${n}_k${kern}_kb${kb}.err | cut -c1-200)"
  done
done
