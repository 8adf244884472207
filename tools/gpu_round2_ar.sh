#!/bin/bash
# round 2, GPU call AR (what is left of the budget; needs profiles/ozaki_batched_epilogue_r2ar.patch applied): batched epilogue of the INT8 kernels -- bit-identity against the serial form in every
# form, model-level checks against the oracle goldens, evaluation time off / on; then as much of the GPU suite as fits with the option ON
mkdir -p gpurun_out
timeout 75 python tools/oz_epi_check.py > gpurun_out/r2ar_check.log 2>&1; echo "check rc=$?"; cat gpurun_out/r2ar_check.log
export GPR_OZ_EPI=1
timeout 120 python -m pytest tests/test_gpu_parity.py -x -v -p no:cacheprovider -k "ozaki or config3_n32768 or mgpu_potrf_on_int8 or config5_n16384 or loss_reference or predict_reference or split_reference" > gpurun_out/r2ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ar_pytest.log
