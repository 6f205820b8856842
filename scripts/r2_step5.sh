#!/bin/bash
# round 2: fp16-scaled gradient planes -- unit test, end-to-end gradient budgets, step time
mkdir -p gpurun_out
LOG=gpurun_out/r2_step5.log
: > $LOG
timeout 600 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k "bn_relu_backward or dgrad or wgrad" 2>&1 | tail -8 >> $LOG
echo "== unit rc $?" >> $LOG
timeout 1200 python -m pytest tests/test_infonce_gpu.py tests/test_cfg2_gpu.py tests/test_r50_gpu.py tests/test_ext_gpu.py -q -p no:cacheprovider 2>&1 | tail -12 >> $LOG
echo "== e2e rc $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.load(open('gpurun_out/test_diag.json'))
for k in sorted(d):
    if ("grad_median" in k or k.startswith("bn_bwd") or k.startswith("precision/") or "logits_vs_fp64" in k or k=="cfg2/grad_exceptions"): print(k, d[k])
PY
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --breakdown --steps 6 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_d.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity", d.get("parity")); print("roofline frac", d["roofline"]["frac"]); print(d["roofline"]["step_breakdown_ms"])
PY
head -40 gpurun_out/r2_bench_d.err >> $LOG
cat $LOG
