#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step14.log
: > $LOG
timeout 400 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k "maxpool or split or bn_" 2>&1 | tail -4 >> $LOG
for sc in "" 1; do
  if [ -n "$sc" ]; then export COCLR_POOL_BWD_SCATTER=1; fi
  echo "---- scatter forced=$sc" >> $LOG
  timeout 120 python tests/tools/run_one_pool.py 64 16 64 64 1 2 1 2>&1 | grep bwd >> $LOG
  timeout 120 python tests/tools/run_one_pool.py 192 16 32 32 1 2 1 2>&1 | grep bwd >> $LOG
  timeout 120 python tests/tools/run_one_pool.py 480 16 16 16 3 2 1 2>&1 | grep bwd >> $LOG
  timeout 120 python tests/tools/run_one_pool.py 832 8 8 8 2 2 0 2>&1 | grep bwd >> $LOG
done
unset COCLR_POOL_BWD_SCATTER
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --no-mixed --breakdown --steps 8 --warmup 3 > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; echo "bench exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_g.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity ok", d["parity"]["ok"], d["parity"]["logits_rel_err"]); print(d["roofline"]["step_breakdown_ms"])
PY
cat $LOG
