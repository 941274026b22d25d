#!/bin/bash
# First hardware run of everything written after round 1's GPU minutes were spent (DESIGN.md sections 0 and 5.7):
# StokesSphericalBEM, the Gauss rules above 4 points, the treecode evaluators of the BEM and Stokes classes,
# fmmb_plan_direct_panels, config C5 at N = 10M.  One gpurun call, about 8 minutes of box time:
#   gpurun --timeout 1200 -- 'bash scripts/gpu_first_run_new_kernels.sh'
# Everything it writes lands in gpurun_out/; copy what should be judged into profiles/.
set -x
mkdir -p gpurun_out
# 1. the guarded suites, without the xfail mask (--runxfail turns xfail marks off: failures show as failures)
timeout 900 python -m pytest tests/test_zz_stokes_bem.py tests/test_zz_bem_rules.py tests/test_zz_c5_full_size.py -q --runxfail \
    2>&1 | tail -60 > gpurun_out/zz_first_run.log
tail -5 gpurun_out/zz_first_run.log
# 2. the drivers: ours (device GMRES), the reference's unchanged driver over the GPU plan, larger sphere
export LD_LIBRARY_PATH=$PWD/fmm_bem_relaxed_b200:$LD_LIBRARY_PATH
( cd gpurun_out && for r in 5 6 7; do
    timeout 300 ../fmm_bem_relaxed_b200/hostcxx/bin/stokes_bem -recursions $r -p 8 -k 4 -solver_tol 1e-5 -check 100
    timeout 300 ../fmm_bem_relaxed_b200/hostcxx/bin/ref_StokesBEM -recursions $r -p 8 -k 4 -solver_tol 1e-5 | grep -v "^P2P"
  done ) > gpurun_out/stokes_bem_drivers.log 2>&1
grep -E "iterations|Fx|solve|setup|matvec vs" gpurun_out/stokes_bem_drivers.log | head -40
# 3. launch list of one solve (per-launch times; share of the near-field and translation kernels)
( cd gpurun_out && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file launches_stokes_bem.csv ../fmm_bem_relaxed_b200/hostcxx/bin/stokes_bem -recursions 7 -p 8 -k 4 -solver_tol 1e-5 \
    > ncu_stokes_bem.log 2>&1 )
# 4. one full capture of the per-matvec near-field kernel (HBM bound: 48 B per pair)
( cd gpurun_out && timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbem_near_kernel -c 1 \
    -o prof_sbem_near ../fmm_bem_relaxed_b200/hostcxx/bin/stokes_bem -recursions 7 -p 8 -k 4 -solver_tol 1e-5 \
    > ncu_sbem_near.log 2>&1 )
ls -la gpurun_out | tail -12
