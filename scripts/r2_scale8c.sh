#!/bin/bash
# 8-GPU line of record with the final code of the round (TMA-staged weight gradient): K=2048, parity block, timeline
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29541 bench.py --gpus 8 --steps 12 --warmup 4 --no-cpu-baseline --no-stock-gpu --no-mixed --timeline > gpurun_out/r2_bench_n8c.json 2> gpurun_out/r2_bench_n8c.err; echo "n8 exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n8c.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s host_ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["config"]["host_enqueue_ms_per_step"]))
print("   parity", d.get("parity")); print("   replicas", d.get("replicas_identical"))
print("   timeline", d["config"].get("phase_timeline_ms")); print("   clocks", d.get("clocks"))
PY
tail -3 gpurun_out/r2_bench_n8c.err
