#!/bin/bash
# round 2, call A (1 GPU): the GPU test suite, then every BASELINE config through bench.py (ours + reference arm)
set -u
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for cfg in disc1m disc16k shipped cluster; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "bench $cfg rc=$?"
  timeout 600 python bench.py --impl reference --config $cfg --steps 3 --warmup 1 > $O/ref_$cfg.json 2> $O/ref_$cfg.err; echo "ref $cfg rc=$?"
done
