python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/conv_stack_run.py 3 1 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stack_v12.csv python tools/conv_stack_run.py 3 1 > gpurun_out/ncu_l2.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v13.log 2>&1; head -c 330 gpurun_out/bench_v13.log; echo; grep -o '"breakdown_ms.*"conv_only' gpurun_out/bench_v13.log
