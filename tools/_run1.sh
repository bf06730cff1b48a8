python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/hier_run.py 4 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hier_v1.csv python tools/hier_run.py 4 > gpurun_out/ncu_l.log 2>&1; cat gpurun_out/plain.log
python tools/conv_stack_run.py 3 1 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stack_v11.csv python tools/conv_stack_run.py 3 1 > gpurun_out/ncu_l2.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v12.log 2>&1; head -c 330 gpurun_out/bench_v12.log; echo; grep -o '"breakdown_ms.*"conv_only' gpurun_out/bench_v12.log
