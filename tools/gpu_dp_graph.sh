#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/dp_graph_exp.py > gpurun_out/dp_graph.log 2>&1; echo "dp graph exp exit $?"; tail -12 gpurun_out/dp_graph.log | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider > gpurun_out/dp_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/dp_tests.log
