python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/dfaust_layer_run.py seg_head 6 1; python tools/dfaust_layer_run.py enc1_block0 6 1; python tools/dfaust_layer_run.py dec2 6 1
python tools/dfaust_layer_run.py seg_head 4 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_seghead_v9.csv python tools/dfaust_layer_run.py seg_head 4 1 > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v9.log 2>&1; tail -c 1800 gpurun_out/bench_v9.log
