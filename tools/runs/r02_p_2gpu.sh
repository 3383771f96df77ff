#!/bin/bash
# round 2, call P (2 GPUs): final check of the sharded paths (warp-level and CTA-level two-sided kernels, one-sided rows, merge,
# the C++ multi-GPU driver) and the 2-GPU bench line
set -u
mkdir -p gpurun_out/r02p
O=gpurun_out/r02p
timeout 420 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu.log
tail -4 $O/pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?"
timeout 100 $TR bench.py --gpus 2 --config cluster --steps 5 --warmup 3 > $O/bench_cluster_2gpu.json 2> $O/bench_cluster_2gpu.err; echo "bench cluster rc=$?"
