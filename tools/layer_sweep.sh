#!/bin/bash
# BASELINE config 1 and a subset of config 5 (single same-level layer, constant density, r = 0.1 -> k-bar ~ 31)
set -u
echo "# config 1: N = 8192, F = 2, 32 -> 64 (fp32 exactness mode, then bf16 tensor-core mode)"
python tools/layer_bench.py --n 8192 --frames 2 --cin 32 --cout 64 --precision 0 --iters 30
python tools/layer_bench.py --n 8192 --frames 2 --cin 32 --cout 64 --precision 1 --iters 30
echo "# config 5 subset (bf16 mode)"
for n in 65536 262144 1048576; do
  for f in 1 2 4; do
    for c in 32 64 128; do
      if [ $((n * f * c)) -gt $((1048576 * 2 * 128)) ]; then continue; fi
      timeout 300 python tools/layer_bench.py --n $n --frames $f --cin $c --cout $c --precision 1 --iters 5 --warmup 2 --density-scale || echo "{\"n\": $n, \"frames\": $f, \"cin\": $c, \"failed\": true}"
    done
  done
done
timeout 300 python tools/layer_bench.py --n 4194304 --frames 1 --cin 32 --cout 32 --precision 1 --iters 3 --warmup 1 --density-scale
