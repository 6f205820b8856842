#!/bin/bash
# round 2, third pass: elect-issued MMAs + stacked weight operand; full GPU suite, A/B timings, step breakdown
mkdir -p gpurun_out
LOG=gpurun_out/r2_step3.log
: > $LOG
timeout 600 python -m pytest tests/test_conv_tma_gpu.py tests/test_conv_gpu.py -q -p no:cacheprovider 2>&1 | tail -12 >> $LOG
echo "== conv suites rc $?" >> $LOG
for tma in 1 0; do
  export COCLR_TMA=$tma
  echo "---- COCLR_TMA=$tma" >> $LOG
  timeout 120 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 120 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 120 python tests/tools/run_one_conv.py dgrad 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py dgrad 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 192 192 3 1 1 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 128 192 1 3 3 32 16 16 16 >> $LOG 2>&1
done
export COCLR_TMA=1
COCLR_TMA_NOSTACK=1 timeout 120 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
STRIDE=2,1,1 COCLR_TMA_NOSTACK=1 timeout 120 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --breakdown --steps 6 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench exit $?" >> $LOG
head -c 400 gpurun_out/r2_bench_b.json >> $LOG; echo >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_b.json') if l.startswith('{')][-1])
print("parity", d.get("parity")); print("replicas", d.get("replicas_identical")); print("roofline frac", d["roofline"]["frac"], "wgrad", d["roofline"]["wgrad"])
PY
head -60 gpurun_out/r2_bench_b.err >> $LOG
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8 >> $LOG
echo "== full gpu suite rc $?" >> $LOG
tail -120 $LOG
