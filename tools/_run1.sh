python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/profile_hierarchy.py > gpurun_out/host_prof.log 2>&1; head -3 gpurun_out/host_prof.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v10.log 2>&1; tail -c 1500 gpurun_out/bench_v10.log
