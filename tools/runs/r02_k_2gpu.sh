#!/bin/bash
# round 2, call K (2 GPUs): every sharded test, the two-sided flow against one GPU, the C++ multi-GPU driver, the bench line
set -u
mkdir -p gpurun_out/r02k
O=gpurun_out/r02k
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu.log
tail -4 $O/pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/mgpu_sym_check.py 131072 3 > $O/mgpu_sym_2_131072.log 2> $O/mgpu_sym_2_131072.err; echo "sym 131072 rc=$?"
timeout 300 $TR tools/mgpu_sym_check.py 1048576 3 > $O/mgpu_sym_2_1m.log 2> $O/mgpu_sym_2_1m.err; echo "sym 1m rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench1 rc=$?"
grep -h replicas $O/*.log | cut -c1-400
