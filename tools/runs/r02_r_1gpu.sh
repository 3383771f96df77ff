#!/bin/bash
# round 2, call R (1 GPU): one mbarrier arrival per warp instead of per thread in the CTA-level two-sided kernel
set -u
mkdir -p gpurun_out/r02r
O=gpurun_out/r02r
timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-whole-run > $O/bench_disc1m.json 2> $O/bench_disc1m.err; echo "bench rc=$?"
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_sided_force_kernel or cell_sorted or resynchronised or deterministic" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
