#!/bin/bash
# One gpurun call: the round-2 ncu evidence (launch list + full captures of the dominant kernels).
# usage (GPU box): bash scripts/gpu_profile_r02.sh <tag>     e.g. r02_v1
set -o pipefail
tag=${1:-r02_v1}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5"
$B > gpurun_out/bench_plain_$tag.json 2> gpurun_out/bench_plain_$tag.err || { echo "plain bench failed"; tail -5 gpurun_out/bench_plain_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5"
ncu --set full --clock-control none --import-source on -k regex:p2p_pair2_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_p2p_$tag $B1 > gpurun_out/ncu_p2p_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trans_gemm_kernel --launch-skip 20 --launch-count 1 -o gpurun_out/prof_m2l_gemm_$tag $B1 > gpurun_out/ncu_gemm_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:m2l_reduce_tma_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_m2l_reduce_$tag $B1 > gpurun_out/ncu_reduce_$tag.log 2>&1
$B1 --m2l-mode 3 > gpurun_out/bench_plain_sweep_$tag.json 2>> gpurun_out/bench_plain_$tag.err && \
ncu --set full --clock-control none --import-source on -k regex:trans_sweep_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_sweep_$tag $B1 --m2l-mode 3 > gpurun_out/ncu_sweep_$tag.log 2>&1
ls -la gpurun_out/*$tag*
