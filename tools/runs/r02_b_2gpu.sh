#!/bin/bash
# round 2, call B (2 GPUs): the sharded tests (BASELINE sizes, several steps per call, threshold crossing), the
# two-sided flow against one GPU at n = 131072 and 1M, and the bench line with its parity block
set -u
mkdir -p gpurun_out/r02b
O=gpurun_out/r02b
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu.log
tail -3 $O/pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/mgpu_sym_check.py 131072 3 > $O/mgpu_sym_2_131072.log 2> $O/mgpu_sym_2_131072.err; echo "sym 131072 rc=$?"
timeout 300 $TR tools/mgpu_sym_check.py 1048576 3 > $O/mgpu_sym_2_1m.log 2> $O/mgpu_sym_2_1m.err; echo "sym 1m rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?"
grep -h replicas $O/*.log | cut -c1-400
