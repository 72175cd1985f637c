set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/bench_final_err.log > gpurun_out/bench_final.json; tail -c 600 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference 2>/dev/null > gpurun_out/bench_final_reference.json; cut -c1-300 gpurun_out/bench_final_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_final.log 2>&1; tail -1 gpurun_out/launches_final.log | cut -c1-200
timeout 600 python tools/bench_kernels.py 2>&1 > gpurun_out/kernels_final.jsonl; wc -l gpurun_out/kernels_final.jsonl
timeout 200 python tools/bench_latency.py --graph 0 2>&1 | grep '^{' > gpurun_out/latency_final.jsonl
timeout 200 python tools/bench_latency.py --graph 1 2>&1 | grep '^{' >> gpurun_out/latency_final.jsonl; cut -c100-300 gpurun_out/latency_final.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 7 -c 1 -o gpurun_out/k2_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/k2_final_ncu.log 2>&1; tail -2 gpurun_out/k2_final_ncu.log
