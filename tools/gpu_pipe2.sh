#!/bin/bash
# 2-GPU pipeline check with hard time limits (a hang must not hold the box)
mkdir -p gpurun_out
HPB_TEST_WATCHDOG=90 NCCL_DEBUG=WARN timeout 150 python -m pytest tests/test_gpu_pipeline.py::test_two_gpu_pipeline_matches_single_gpu -x -q -s > gpurun_out/pipe2.log 2>&1
echo "rc=$?"; grep -E "rank [01]\]|passed|failed|Error|File|line |NCCL WARN" gpurun_out/pipe2.log | tail -25
if grep -q "1 passed" gpurun_out/pipe2.log; then
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 2 --warmup 3 2>gpurun_out/bench_2gpu.err | tee gpurun_out/bench_2gpu.json
  tail -5 gpurun_out/bench_2gpu.err
fi
