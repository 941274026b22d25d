#!/bin/bash
# usage (on the GPU box via gpurun): bash scripts/gpu_tests.sh
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tee gpurun_out/pytest_gpu.log
