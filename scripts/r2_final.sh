#!/bin/bash
# what the driver runs at round end (GPU tests, smoke, default bench) plus the other bench arms of this round
mkdir -p gpurun_out
LOG=gpurun_out/r2_final.log
: > $LOG
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6 >> $LOG
echo "== gpu suite rc $?" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -3 >> $LOG
timeout 900 python bench.py --timeline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $?" >> $LOG
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference arm exit $?" >> $LOG
timeout 600 python bench.py --model coclr --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_coclr.json 2> gpurun_out/r2_bench_coclr.err; echo "coclr exit $?" >> $LOG
timeout 600 python bench.py --net r50 --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-mixed > gpurun_out/r2_bench_r50.json 2> gpurun_out/r2_bench_r50.err; echo "r50 exit $?" >> $LOG
timeout 600 python bench.py --moco-k 16384 --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-mixed > gpurun_out/r2_bench_k16384.json 2> gpurun_out/r2_bench_k16384.err; echo "k16384 exit $?" >> $LOG
timeout 600 python bench.py --net s3dg --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-mixed > gpurun_out/r2_bench_s3dg.json 2> gpurun_out/r2_bench_s3dg.err; echo "s3dg exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
for f in ("r2_bench_default","r2_bench_reference","r2_bench_coclr","r2_bench_r50","r2_bench_k16384","r2_bench_s3dg"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, d["metric"], "value %.1f ms %.2f" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e") and round(d["e2e"]["value"],1), "launches", d.get("gpu_launches"))
        for k in ("parity","replicas_identical","stock_gpu_baseline","cpu_baseline","clocks"): 
            if d.get(k) is not None: print("   ",k, d[k])
        if "roofline" in d: print("    roofline", {k:v for k,v in d["roofline"].items() if k in ("achieved","frac","traffic","ms_per_step_in_kernel","wgrad","step_breakdown_ms")})
        c=d.get("config",{})
        for k in ("mixed_precision","phase_timeline_ms","host_enqueue_ms_per_step"): 
            if c.get(k) is not None: print("   ",k,c[k])
    except Exception as ex:
        print(f, "failed", repr(ex)); 
        try: print(open('gpurun_out/%s.err'%f).read()[-1500:])
        except Exception: pass
PY
cat $LOG
