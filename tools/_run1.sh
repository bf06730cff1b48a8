python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/hier_run.py 6 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v14.log 2>&1; head -c 330 gpurun_out/bench_v14.log; echo; grep -o '"breakdown_ms.*"conv_only' gpurun_out/bench_v14.log
