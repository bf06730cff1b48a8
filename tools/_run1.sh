python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v16.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v16.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['breakdown_ms'])
r=d['roofline']; print(r['kernel'], r['frac'], r['kernel_ms'])
for k in r['gather_kernels_per_step']: print(k)
for k in r['largest_launch_all_kernels']: print(k)
PY
