#!/bin/bash
# 2-GPU visit: NCCL pipeline parity (beam and laser packets) and the N = 2 bench line with its verification leg
tag=${1:-r02h}
mkdir -p gpurun_out
HPB_TEST_WATCHDOG=120 timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -rfEs --tb=short --timeout 300 2>&1 | tail -30 > gpurun_out/${tag}_pytest.txt
tail -5 gpurun_out/${tag}_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/${tag}_bench2.err | tee gpurun_out/${tag}_bench2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>> gpurun_out/${tag}_bench2.err | tee gpurun_out/${tag}_bench2_reference.json
tail -3 gpurun_out/${tag}_bench2.err
