#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2q_pytest.log
timeout 900 python experiments/sample_cluster/run_gpu.py > gpurun_out/r2q_sample_cluster.log 2>&1; echo "cluster sampler rc=$?"; tail -4 gpurun_out/r2q_sample_cluster.log
