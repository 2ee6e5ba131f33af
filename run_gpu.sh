#!/bin/bash
# scratch: GPU run 12
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/run12_tests.log 2>&1
tail -5 gpurun_out/run12_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/run12_bench.log 2>&1
tail -1 gpurun_out/run12_bench.log
