#!/bin/bash
# bench lines of record with the final code of the round (the full GPU suite ran green in scripts/r2_final.sh / r2_step21.sh)
mkdir -p gpurun_out
timeout 400 python bench.py --timeline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $?"
timeout 200 python bench.py --model coclr --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-mixed > gpurun_out/r2_bench_coclr.json 2> gpurun_out/r2_bench_coclr.err; echo "coclr exit $?"
timeout 200 python bench.py --net r50 --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-mixed > gpurun_out/r2_bench_r50.json 2> gpurun_out/r2_bench_r50.err; echo "r50 exit $?"
timeout 200 python bench.py --moco-k 16384 --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-mixed > gpurun_out/r2_bench_k16384.json 2> gpurun_out/r2_bench_k16384.err; echo "k16384 exit $?"
timeout 200 python bench.py --net s3dg --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-mixed > gpurun_out/r2_bench_s3dg.json 2> gpurun_out/r2_bench_s3dg.err; echo "s3dg exit $?"
python - <<'PY'
import json
for f in ("r2_bench_default","r2_bench_coclr","r2_bench_r50","r2_bench_k16384","r2_bench_s3dg"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "value %.1f ms %.2f" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e") and round(d["e2e"]["value"],1), "launches", d.get("gpu_launches"), "parity", (d.get("parity") or {}).get("ok"), "frac", d["roofline"]["frac"], "mixed", (d["config"].get("mixed_precision") or {}).get("value"), "clocks", d.get("clocks"))
    except Exception as ex:
        print(f, "failed", repr(ex))
PY
