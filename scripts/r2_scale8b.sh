#!/bin/bash
# 8-GPU lines of record (final code of the round): K=2048 with the parity block + timeline, config 3 (K=16384)
mkdir -p gpurun_out
LOG=gpurun_out/r2_scale8b.log
: > $LOG
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $R --master-port 29531 bench.py --gpus 8 --steps 12 --warmup 4 --no-cpu-baseline --no-stock-gpu --no-mixed --timeline > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err; echo "n8 exit $?" >> $LOG
timeout 400 $R --master-port 29533 bench.py --gpus 8 --steps 12 --warmup 4 --no-cpu-baseline --no-stock-gpu --no-mixed --no-parity --moco-k 16384 > gpurun_out/r2_bench_n8b_k16384.json 2> gpurun_out/r2_bench_n8b_k16384.err; echo "n8 K=16384 exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
for f in ("r2_bench_n8b","r2_bench_n8b_k16384"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "value %.0f ms %.2f e2e %s host_ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["config"]["host_enqueue_ms_per_step"]))
        print("   parity", d.get("parity")); print("   replicas", d.get("replicas_identical"))
        print("   timeline", d["config"].get("phase_timeline_ms")); print("   clocks", d.get("clocks"))
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -3 gpurun_out/r2_bench_n8b.err >> $LOG
cat $LOG
