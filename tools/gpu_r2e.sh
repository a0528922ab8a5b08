#!/bin/bash
TAG=r2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_batched_options_gpu.py -x -q --durations=8 > gpurun_out/${TAG}_options_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_options_tests.log
tail -30 gpurun_out/${TAG}_options_tests.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_batched_options_gpu.py > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 30 --no-cpu --also standard > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
