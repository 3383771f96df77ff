#!/bin/bash
# round 2, call C (1 GPU): the rewritten two-sided kernel (fixed-point sums, sparse exact redo): parity tests, then speed
set -u
mkdir -p gpurun_out/r02c
O=gpurun_out/r02c
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log
for cfg in disc1m cluster disc16k shipped; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "bench $cfg rc=$?"
done
timeout 300 python tools/config_runs.py cluster 200 > $O/run_cluster.json 2> $O/run_cluster.err
