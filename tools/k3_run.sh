#!/bin/bash
# K3: parity of the sparse head (padded + packed), then c4head bench lines at B = 16 / 64 / 256 documents
mkdir -p gpurun_out/k3
timeout 900 python -m pytest tests -m gpu -q -x -k "sparse_head or quantis or topk_sampling or smoke" 2>&1 | tail -4
for B in 16 64 256; do
  timeout 600 python bench.py --config c4head --queries $B --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k3/c4head_b$B.json 2> gpurun_out/k3/c4head_b$B.err
  echo "B=$B rc=$? $(python -c "
import json
d=json.load(open('gpurun_out/k3/c4head_b$B.json'))
print(round(d['ms_per_step'],3),'ms; gemm',round(d['roofline']['kernel_ms'],3),'ms frac',round(d['roofline']['frac'],4),'share',round(d['roofline']['kernel_share_of_step'],3),'parity',d['parity']['ok'], 'e2e', round(d['e2e']['value'],1), d['unit'])
" 2>&1 | tail -1) $(tail -1 gpurun_out/k3/c4head_b$B.err | cut -c1-200)"
done
