#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step21.log
: > $LOG
timeout 200 python tests/tools/smoke_layerwise.py 2>&1 | grep -v "Warn\|Consider\|float(" > gpurun_out/smoke_layers_new.txt
head -1 gpurun_out/smoke_layers_new.txt >> $LOG
grep "Mixed_5b.branch3\|Mixed_5c.branch1.1.bn2\|Mixed_4f.branch3\|Conv_2c.bn2 \|<<<" gpurun_out/smoke_layers_new.txt | head -8 >> $LOG
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke >> $LOG
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_conv_tma_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -3 >> $LOG
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --no-mixed --steps 8 --warmup 3 > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err; echo "bench exit $?" >> $LOG
python - >> $LOG <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_k.json') if l.startswith('{')][-1])
print("value %.0f ms %.2f e2e %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), d["gpu_launches"]))
print("parity", d["parity"]["ok"], d["parity"]["logits_rel_err"], d["parity"]["queue_rel_err"]); print(d["roofline"]["frac"], d["roofline"]["step_breakdown_ms"]["coclr_conv_igemm"])
PY
timeout 900 python -m pytest tests/test_infonce_gpu.py tests/test_cfg2_gpu.py tests/test_s3dg_gpu.py tests/test_r50_gpu.py tests/test_ext_gpu.py -q -p no:cacheprovider 2>&1 | tail -4 >> $LOG
python - >> $LOG <<'PY'
import json
d=json.load(open('gpurun_out/test_diag.json'))
for k in ('infonce/grad_median_new_ref','s3dg/grad_median_new_ref','cfg2/grad_median_new_ref','r50/grad_median_new_ref','infonce/layerwise_max_rel_l2','infonce/logits_vs_fp64','s3dg/logits_vs_fp64'):
    if k in d: print(k, d[k])
PY
cat $LOG
