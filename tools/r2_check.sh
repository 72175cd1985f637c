#!/bin/bash
# round-2 re-entry check: GPU suite, smoke, a few bench configs at shard sizes, default bench
mkdir -p gpurun_out/chk
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in "c4 --docs 1100000" "c4uniform --docs 1100000" "c3m128 --docs 1100000" "c3m256 --docs 1100000" "c2k1000 --docs 1100000" "c1"; do
  n=$(echo $c | cut -d' ' -f1)
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/chk/$n.json 2> gpurun_out/chk/$n.err
  echo "$n rc=$? $(cut -c1-300 gpurun_out/chk/$n.json) $(tail -2 gpurun_out/chk/$n.err | cut -c1-300)"
done
timeout 600 python tools/bench_kernels.py 2>&1 > gpurun_out/chk/kernels.jsonl; wc -l gpurun_out/chk/kernels.jsonl
