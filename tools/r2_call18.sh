#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q > gpurun_out/r2c18_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c18_tests.log
timeout 300 python tools/attn_prof.py 2443 64 > gpurun_out/r2c18_attn.txt 2>&1
timeout 300 python tools/attn_prof.py 2560 64 >> gpurun_out/r2c18_attn.txt 2>&1
timeout 300 python tools/attn_prof.py 1370 192 >> gpurun_out/r2c18_attn.txt 2>&1
timeout 600 python -m pytest tests/test_depth_gpu.py -m gpu -q >> gpurun_out/r2c18_tests.log 2>&1
echo "depth tests rc=$?" >> gpurun_out/r2c18_tests.log
tail -6 gpurun_out/r2c18_tests.log; cat gpurun_out/r2c18_attn.txt
