#!/bin/bash
# quick functional pass over every bench config at reduced sizes (development aid)
mkdir -p gpurun_out/sb
for c in "c1" "c2 --docs 1100000" "c3m128 --docs 1100000" "c5b32 --docs 1100000" "c5b1 --docs 1100000" "c2k1000 --docs 1100000" "c4uniform --docs 1100000" "c4 --docs 1100000" "c4head --queries 16"; do
  n=$(echo $c | cut -d' ' -f1)
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sb/$n.json 2> gpurun_out/sb/$n.err
  echo "$n rc=$? $(cut -c1-200 gpurun_out/sb/$n.json) $(tail -2 gpurun_out/sb/$n.err | cut -c1-400)"
done
