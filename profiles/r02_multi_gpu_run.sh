#!/bin/bash
# Round-2 multi-GPU validation on one 8 x B200 box (outputs under gpurun_out/): bench.py at N GPUs (headline + e2e legs only), the
# two-device / relay / group tests, the C++ group pipeline on all GPUs.   usage: bash profiles/r02_multi_gpu_run.sh "8 2"
cd "$(dirname "$0")/.."
for N in ${1:-8}; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 10 --warmup 3 --skip-ops --skip-cpu \
    > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
  echo "bench N=$N rc=$? $(head -c 150 gpurun_out/r02_bench_n$N.json)"
done
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout=150 -k "two_devices or relay or cpp_group_pipeline" > gpurun_out/r02_multi_tests.txt 2>&1
tail -3 gpurun_out/r02_multi_tests.txt
rm -f gpurun_out/r02_group_pipeline.txt
g++ -std=c++17 -O2 -I include tests/cpp/group_pipeline.cpp -L pvac_hfhe_cppbyv_b200 -lpvacb -Wl,-rpath,$PWD/pvac_hfhe_cppbyv_b200 -o /tmp/group_pipeline \
  && for N in 1 8; do timeout 60 /tmp/group_pipeline 524288 32768 $N >> gpurun_out/r02_group_pipeline.txt 2>&1; done
cat gpurun_out/r02_group_pipeline.txt
