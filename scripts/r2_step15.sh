#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step15.log
: > $LOG
timeout 500 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k "wgrad" 2>&1 | tail -4 >> $LOG
STRIDE=2,1,1 SPLITS=148 timeout 60 python tests/tools/run_one_conv.py wgrad 64 64 7 1 1 32 32 64 64 2>&1 | tail -1 >> $LOG
SPLITS=74 timeout 60 python tests/tools/run_one_conv.py wgrad 192 64 1 1 1 32 16 16 16 2>&1 | tail -1 >> $LOG
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --no-mixed --breakdown --steps 8 --warmup 3 > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err; echo "bench exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_h.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity ok", d["parity"]["ok"], d["parity"]["logits_rel_err"]); print(d["roofline"]["step_breakdown_ms"])
PY
timeout 600 python -m pytest tests/test_infonce_gpu.py tests/test_cfg2_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -4 >> $LOG
cat $LOG
