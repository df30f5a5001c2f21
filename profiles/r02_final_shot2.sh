#!/bin/bash
# 27 s of box time left: the small reference programs and test_depth on the GPU, two pytest processes side by side
mkdir -p gpurun_out/final
export PVACB_SAVE_OUTPUT=$PWD/gpurun_out/final
(timeout 23 python -m pytest tests/test_gpu_round2.py -q -p no:cacheprovider -k "reference_small" > gpurun_out/final/small.txt 2>&1; echo "rc=$?" >> gpurun_out/final/small.txt) &
(timeout 23 python -m pytest tests/test_gpu_round2.py -q -p no:cacheprovider -k "reference_test_depth" > gpurun_out/final/depth.txt 2>&1; echo "rc=$?" >> gpurun_out/final/depth.txt) &
wait
tail -4 gpurun_out/final/small.txt gpurun_out/final/depth.txt
