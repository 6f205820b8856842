#!/bin/bash
# parity tests + default bench (no ncu)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
cp gpurun_out/test_diag.json gpurun_out/test_diag_full.json 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --breakdown > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None)
print(d["roofline"]["step_breakdown_ms"])
PY
head -40 gpurun_out/bench_quick.err | tail -28
