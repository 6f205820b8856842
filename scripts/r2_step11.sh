#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step11.log
: > $LOG
timeout 300 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k "maxpool" 2>&1 | tail -5 >> $LOG
echo "== pool tests rc $?" >> $LOG
for reg in "" 1; do
  if [ -n "$reg" ]; then export COCLR_POOL133_REG=1; fi
  echo "---- 133 register-only kernel=$reg" >> $LOG
  timeout 120 python tests/tools/run_one_pool.py 64 16 64 64 1 2 1 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_pool.py 192 16 32 32 1 2 1 >> $LOG 2>&1
done
unset COCLR_POOL133_REG
for sp in 24 37 49 74 98; do
  SPLITS=$sp timeout 60 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 2>&1 | sed "s/^/splits=$sp /" >> $LOG
done
for sp in 74 148 296; do
  STRIDE=2,1,1 SPLITS=$sp timeout 60 python tests/tools/run_one_conv.py wgrad 64 64 7 1 1 32 32 64 64 2>&1 | sed "s/^/splits=$sp /" >> $LOG
done
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --breakdown --steps 8 --warmup 3 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "bench exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_f.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity ok", d["parity"]["ok"], d["parity"]["logits_rel_err"]); print("roofline frac", d["roofline"]["frac"]); print(d["roofline"]["step_breakdown_ms"]); print(d["config"]["mixed_precision"])
PY
grep "launches with M" gpurun_out/r2_bench_f.err >> $LOG
cat $LOG
