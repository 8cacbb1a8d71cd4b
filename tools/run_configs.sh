#!/bin/bash
# Measures the BASELINE.json configs that are parity-test cases rather than the bench line
# (configs[1], [3], and a per-GPU slice of [4]) on ONE B200.  Usage: tools/run_configs.sh > gpurun_out/configs.jsonl
set -u
cd "$(dirname "$0")/.."
run() { echo "## $*" >&2; python bench.py --no-cpu-baseline --extra-batches "" "$@" | tail -1; }
# config 2: 1M x 384 fp32 cosine top-10, B = 1 / 32 / 1024
for B in 1 32 1024; do run --rows 1000000 --dim 384 --dtype f32 --batch $B --steps 100 --warmup 5; done
# config 4: 10M x 384 with `where` at 1 / 10 / 50 % + 5 % tombstones (bf16 and fp32 rows), B = 1
for S in 0.01 0.1 0.5; do run --rows 10000000 --dim 384 --dtype bf16 --batch 1 --selectivity $S --tombstones 0.05 --steps 100 --warmup 5; done
for S in 0.01 0.1 0.5; do run --rows 10000000 --dim 384 --dtype f32 --batch 1 --selectivity $S --tombstones 0.05 --steps 100 --warmup 5; done
run --rows 10000000 --dim 384 --dtype bf16 --batch 1 --steps 100 --warmup 5
run --rows 10000000 --dim 384 --dtype f32 --batch 1 --steps 100 --warmup 5
# config 5, one GPU's share: 25M x 384 bf16, B = 1024, top-100, l2
run --rows 25000000 --dim 384 --dtype bf16 --space l2 --batch 1024 --k 100 --steps 10 --warmup 3
run --rows 25000000 --dim 384 --dtype bf16 --space l2 --batch 1024 --k 10 --steps 10 --warmup 3
# headline shape with k = 100
run --batch 1024 --k 100 --steps 10 --warmup 3
