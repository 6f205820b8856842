#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step18.log
: > $LOG
timeout 400 python -m pytest tests/test_wgrad_tma_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -15 >> $LOG
echo "== wgrad_tma rc ${PIPESTATUS[0]}" >> $LOG
for ws in 1 ""; do
  echo "--- WS=$ws" >> $LOG
  WS=$ws timeout 60 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 2>&1 | tail -1 >> $LOG
  WS=$ws timeout 60 python tests/tools/run_one_conv.py wgrad 192 192 3 1 1 32 16 32 32 2>&1 | tail -1 >> $LOG
  WS=$ws timeout 60 python tests/tools/run_one_conv.py wgrad 256 128 1 1 1 32 16 16 16 2>&1 | tail -1 >> $LOG
  WS=$ws timeout 60 python tests/tools/run_one_conv.py wgrad 112 224 1 3 3 32 8 8 8 2>&1 | tail -1 >> $LOG
  WS=$ws timeout 60 python tests/tools/run_one_conv.py wgrad 832 384 1 1 1 32 4 4 4 2>&1 | tail -1 >> $LOG
  WS=$ws timeout 60 python tests/tools/run_one_conv.py wgrad 64 64 1 1 1 32 16 32 32 2>&1 | tail -1 >> $LOG
done
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --no-mixed --breakdown --steps 8 --warmup 3 > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; echo "bench exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_i.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity ok", d["parity"]["ok"], d["parity"]["logits_rel_err"]); print(d["roofline"]["step_breakdown_ms"])
PY
tail -5 gpurun_out/r2_bench_i.err >> $LOG
timeout 900 python -m pytest tests/test_infonce_gpu.py tests/test_cfg2_gpu.py tests/test_r50_gpu.py tests/test_s3dg_gpu.py -q -p no:cacheprovider 2>&1 | tail -6 >> $LOG
cat $LOG
