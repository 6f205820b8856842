#!/bin/bash
# 2-GPU pass: NCCL parity test, bench parity block at N=2, overlapped gradient all-reduce, mask_topk / batch packing
mkdir -p gpurun_out
LOG=gpurun_out/r2_dist2.log
: > $LOG
timeout 900 python -m pytest tests/test_ext_gpu.py tests/test_dist_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -8 >> $LOG
echo "== ext+dist rc $?" >> $LOG
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --steps 6 --warmup 3 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; echo "bench1 exit $?" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "bench2 exit $?" >> $LOG
COCLR_OVERLAP_ALLREDUCE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_c2_noov.json 2> gpurun_out/r2_bench_c2_noov.err; echo "bench2 (no overlap) exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
for f in ("r2_bench_c1","r2_bench_c2","r2_bench_c2_noov"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "value %.0f ms %.2f e2e %s launches %s host_ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"], d["config"]["host_enqueue_ms_per_step"]))
        print("   parity", d.get("parity")); print("   replicas", d.get("replicas_identical"))
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -5 gpurun_out/r2_bench_c2.err >> $LOG
cat $LOG
