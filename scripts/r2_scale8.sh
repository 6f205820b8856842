#!/bin/bash
# 8-GPU session: K=2048 with and without the overlapped all-reduce (+ per-phase timeline), config 3 (K=16384)
mkdir -p gpurun_out
LOG=gpurun_out/r2_scale8.log
: > $LOG
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 4 --no-cpu-baseline --timeline > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "n8 exit $?" >> $LOG
COCLR_OVERLAP_ALLREDUCE=0 timeout 600 $R --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 4 --no-cpu-baseline --no-parity --no-mixed --no-e2e --timeline > gpurun_out/r2_bench_n8_noov.json 2> gpurun_out/r2_bench_n8_noov.err; echo "n8 no-overlap exit $?" >> $LOG
timeout 600 $R --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 4 --no-cpu-baseline --moco-k 16384 --no-mixed > gpurun_out/r2_bench_n8_k16384.json 2> gpurun_out/r2_bench_n8_k16384.err; echo "n8 K=16384 exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
for f in ("r2_bench_n8","r2_bench_n8_noov","r2_bench_n8_k16384"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "value %.0f ms %.2f e2e %s host_ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["config"]["host_enqueue_ms_per_step"]))
        print("   parity", d.get("parity")); print("   replicas", d.get("replicas_identical")); print("   mixed", d["config"].get("mixed_precision"))
        print("   timeline", d["config"].get("phase_timeline_ms"))
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -3 gpurun_out/r2_bench_n8.err >> $LOG
cat $LOG
