#!/bin/bash
# round 2, GPU call AP (the last 4 GPU-minutes): cluster pairs with a multicast op(B) tile for the 128 x 128 window kernels of the INT8 route
# (option ozaki_mc): A/B of the isolated products and of the N = 32768 evaluation, then the whole GPU suite with the option ON
mkdir -p gpurun_out
timeout 110 python tools/oz_mc_ab.py > gpurun_out/r2ap_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2ap_ab.log | tail -20
export GPR_OZ_MC=1
timeout 200 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2ap_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2ap_pytest.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/r2ap_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2ap_smoke.log
