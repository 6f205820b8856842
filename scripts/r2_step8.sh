#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step8.log
: > $LOG
timeout 600 python -m pytest tests/test_conv_tma_gpu.py tests/test_conv_gpu.py -q -p no:cacheprovider 2>&1 | tail -5 >> $LOG
echo "== conv suites rc $?" >> $LOG
timeout 60 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
timeout 60 python tests/tools/run_one_conv.py fwd 64 256 1 1 1 32 16 32 32 >> $LOG 2>&1
timeout 60 python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --breakdown --steps 6 --warmup 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_e.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity ok", d["parity"]["ok"], d["parity"]["logits_rel_err"]); print("roofline frac", d["roofline"]["frac"]); print(d["roofline"]["step_breakdown_ms"])
PY
sed -n 1,14p gpurun_out/r2_bench_e.err >> $LOG
cat $LOG
