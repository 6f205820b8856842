#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step19.log
: > $LOG
COCLR_TMA_DEBUG=1 timeout 120 python -m pytest tests/test_wgrad_tma_gpu.py -q -p no:cacheprovider -x -k "stride2 or window" 2>&1 | tail -25 >> $LOG
echo "== new kinds rc ${PIPESTATUS[0]}" >> $LOG
if ! grep -q "failed\|rror" $LOG; then
  for ws in 1 ""; do
    echo "--- WS=$ws" >> $LOG
    WS=$ws STRIDE=2,1,1 timeout 60 python tests/tools/run_one_conv.py wgrad 64 64 7 1 1 32 32 64 64 2>&1 | tail -1 >> $LOG
  done
  timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --no-mixed --breakdown --steps 8 --warmup 3 > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err; echo "bench exit $?" >> $LOG
  python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_j.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity ok", d["parity"]["ok"], d["parity"]["logits_rel_err"]); print(d["roofline"]["step_breakdown_ms"])
PY
  grep "conv_wgrad" gpurun_out/r2_bench_j.err | head -8 >> $LOG
  timeout 900 python -m pytest tests/test_wgrad_tma_gpu.py tests/test_infonce_gpu.py tests/test_cfg2_gpu.py -q -p no:cacheprovider 2>&1 | tail -4 >> $LOG
fi
cat $LOG
