#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_infonce_gpu.py -q -m gpu --timeout 600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_infonce.log
echo "== infonce exit ${PIPESTATUS[0]}"; tail -40 gpurun_out/pytest_infonce.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -15
