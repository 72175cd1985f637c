#!/bin/bash
# End-of-round-2 evidence set on ONE GPU (about 12 minutes).  Everything lands in gpurun_out/final/; the summaries that
# are judged are copied to profiles/r2/ by hand (tools/ncu_digest.py turns the .ncu-rep files into text).
O=gpurun_out/final
mkdir -p $O
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 400 $O/bench_c2.json
timeout 600 python bench.py --impl reference > $O/bench_c2_reference.json 2>/dev/null; cut -c1-300 $O/bench_c2_reference.json
for c in c1 c2k1000 c3m128 c3m256 c3m512 c3m1024 c3m128c c3m256c c4 c4uniform c4head c5b1 c5b32; do
  timeout 900 python bench.py --config $c --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err
  echo "$c rc=$? $(cut -c1-220 $O/bench_$c.json)"
done
timeout 600 python bench.py --config c4head --queries 256 --no-cpu-baseline > $O/bench_c4head_b256.json 2>/dev/null
timeout 600 python bench.py --config c4 --docs 1100000 --no-cpu-baseline > $O/bench_c4_1p1M.json 2>/dev/null
timeout 600 python bench.py --config c4 --docs 1100000 --queries 32 --no-cpu-baseline > $O/bench_c4_1p1M_q32.json 2>/dev/null
timeout 600 python bench.py --config c4uniform --docs 1100000 --no-cpu-baseline > $O/bench_c4uniform_1p1M.json 2>/dev/null
timeout 600 python tools/bench_kernels.py k1 k3 > $O/kernels.jsonl 2>/dev/null; wc -l $O/kernels.jsonl
# launch list of the headline step, then full captures of the dominant kernel of each path
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 7 -c 1 -o $O/k2_c2_main python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_k2.log 2>&1; tail -1 $O/ncu_k2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 3 -c 1 -o $O/k3_packed_b64 python bench.py --config c4head --queries 64 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_k3.log 2>&1; tail -1 $O/ncu_k3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_score_rows -s 2 -c 1 -o $O/k4_rows_zipf python bench.py --config c4 --docs 1100000 --queries 2000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_k4.log 2>&1; tail -1 $O/ncu_k4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 19 -c 1 -o $O/k2_mrl128_main python bench.py --config c3m128 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_mrl.log 2>&1; tail -1 $O/ncu_mrl.log
