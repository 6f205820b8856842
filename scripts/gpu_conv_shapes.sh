#!/bin/bash
# micro-timings of representative launch shapes, then parity tests of the conv kernels
STRIDE=2,1,1 python tests/tools/run_one_conv.py dgrad 64 64 7 1 1 32 32 64 64
STRIDE=2,1,1 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64
python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32
python tests/tools/run_one_conv.py dgrad 64 192 1 3 3 32 16 32 32
python tests/tools/run_one_conv.py fwd 64 256 1 1 1 32 16 32 32
python tests/tools/run_one_conv.py fwd 256 64 1 1 1 32 16 32 32
python tests/tools/run_one_conv.py fwd 512 256 1 3 3 32 8 16 16
python tests/tools/run_one_conv.py fwd 480 192 1 1 1 32 8 8 8
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
