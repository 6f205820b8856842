#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step13.log
: > $LOG
timeout 600 python -m pytest tests/test_s3dg_gpu.py -q -p no:cacheprovider 2>&1 | tail -25 >> $LOG
python - >> $LOG <<'PY'
import json, numpy as np
d=json.load(open('gpurun_out/test_diag.json'))
for k in d:
    if k.startswith('s3dg/') and 'grad_err' not in k and 'kernels' not in k: print(k, d[k])
g=d.get('s3dg/grad_err_new_vs_ref', {})
items=sorted(g.items(), key=lambda kv:-kv[1][0])
for k,v in items[:8]: print(k, '%.2e %.2e'%tuple(v))
PY
cat $LOG
