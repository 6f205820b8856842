#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python bench.py --steps 4 --warmup 3 --breakdown > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "== bench exit $?"; cat gpurun_out/bench1.json; tail -60 gpurun_out/bench1.err
for m in tf32 fp32 bf16; do timeout 600 python tests/tools/bench_torch_gpu.py --mode $m --steps 3 2>&1 | tail -2 | tee -a gpurun_out/torch_gpu.jsonl; done
timeout 900 python -m pytest tests/test_infonce_gpu.py -q -m gpu --timeout 600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -15
