#!/bin/bash
# conv/pool unit parity, InfoNCE step parity, then a short bench with the per-kernel breakdown
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --timeout 120 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -25 > gpurun_out/pytest_conv.log
echo "== conv exit ${PIPESTATUS[0]}"; tail -12 gpurun_out/pytest_conv.log
timeout 900 python -m pytest tests/test_infonce_gpu.py -q -m gpu --timeout 600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_infonce.log
echo "== infonce exit ${PIPESTATUS[0]}"; tail -25 gpurun_out/pytest_infonce.log
for prec in parity mixed; do
timeout 600 python bench.py --steps 4 --warmup 3 --breakdown --no-cpu-baseline --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
echo "== bench $prec exit $?"; cat gpurun_out/bench_$prec.json; grep -v Warning gpurun_out/bench_$prec.err | head -45
done
