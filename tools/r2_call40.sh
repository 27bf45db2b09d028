#!/bin/bash
# compute-sanitizer memcheck over the kernels added in the second session: transposed tiles (also with 16 epilogue warps), the
# TMA-store epilogue with bias / GELU, the TMA reduce-add epilogue (small shapes), the merged correlation pyramid + lookup, and
# the fp32-class SOLOv2 backbone (test-size twin)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gemm_gpu.py -m gpu -q \
  -k "256-128-64-0-640 or 777-128-200-1-640 or 4100-68-128-0-640 or 5000-96-576-2-640 or 700-328-200 or 300-264-200 or 2443-384-384" > gpurun_out/r2c40_sanitizer_gemm.log 2>&1
echo "sanitizer gemm rc=$?" >> gpurun_out/r2c40_sanitizer_gemm.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_flow_gpu.py -m gpu -q -k "corr_pyramid" > gpurun_out/r2c40_sanitizer_corr.log 2>&1
echo "sanitizer corr rc=$?" >> gpurun_out/r2c40_sanitizer_corr.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_mask_gpu.py -m gpu -q -k "tiny_exact" > gpurun_out/r2c40_sanitizer_mask.log 2>&1
echo "sanitizer mask rc=$?" >> gpurun_out/r2c40_sanitizer_mask.log
for f in gemm corr mask; do tail -5 gpurun_out/r2c40_sanitizer_$f.log; done
