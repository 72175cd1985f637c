#!/bin/bash
# 8-GPU evidence lines (one node, NCCL): headline c2, sparse c4 (Zipf) through the sharded index, online c5 batch 32 / 1.
# Every run has a SHORT timeout: an 8-GPU call is charged 8x, and a torchrun that hangs in teardown (seen once with the
# online configs, fixed in bench.py) would otherwise burn the round's GPU budget until the call's own limit.
N=${N:-8}
mkdir -p gpurun_out/n$N
for c in ${CONFIGS:-c2 c4 c5b32 c5b1}; do
  n=$(echo $c | cut -d' ' -f1)
  timeout -k 10 ${RUN_TIMEOUT:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/n$N/$n.json 2> gpurun_out/n$N/$n.err
  echo "$n N=$N rc=$? $(grep '^{' gpurun_out/n$N/$n.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value'],1), d['unit'], round(d['ms_per_step'],3),'ms p50', d.get('p50_ms_per_step'), 'frac',round(r['frac'],3),'share',round(r.get('kernel_share_of_step',0),3),'parity',d['parity']['ok'], 'e2e', round(d['e2e']['value'],1), {k:v for k,v in d.items() if 'lat' in k or 'p99' in k})
" 2>&1 | tail -1) $(tail -1 gpurun_out/n$N/$n.err | cut -c1-200)"
done
