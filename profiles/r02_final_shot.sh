#!/bin/bash
# The round's last GPU seconds (125 s of box time were left): the reference's own programs on the GPU, then a short bench.py run.
mkdir -p gpurun_out/final
export PVACB_SAVE_OUTPUT=$PWD/gpurun_out/final
timeout 75 python -m pytest tests/test_gpu_round2.py -q -p no:cacheprovider -k "reference_test_main or reference_small or reference_test_depth" --durations=5 > gpurun_out/final/refprogs.txt 2>&1
echo "rc=$?" >> gpurun_out/final/refprogs.txt
timeout 40 python bench.py --steps 3 --warmup 3 --skip-ops --skip-cpu > gpurun_out/final/bench_quick.json 2> gpurun_out/final/bench_quick.err
echo "rc=$?" >> gpurun_out/final/bench_quick.err
tail -12 gpurun_out/final/refprogs.txt
tail -c 400 gpurun_out/final/bench_quick.json
tail -3 gpurun_out/final/bench_quick.err
