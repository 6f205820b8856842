#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step12.log
: > $LOG
timeout 600 python -m pytest tests/test_s3dg_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -25 >> $LOG
echo "== s3dg tests rc $?" >> $LOG
timeout 600 python bench.py --net s3dg --no-cpu-baseline --no-stock-gpu --no-mixed --steps 6 --warmup 3 > gpurun_out/r2_bench_s3dg.json 2> gpurun_out/r2_bench_s3dg.err; echo "bench exit $?" >> $LOG
tail -3 gpurun_out/r2_bench_s3dg.err >> $LOG
python - >> $LOG <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_s3dg.json') if l.startswith('{')][-1])
    print("s3dg value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
    print("parity", d["parity"])
except Exception as e:
    print("no bench line", e)
PY
cat $LOG
