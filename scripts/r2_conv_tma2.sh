#!/bin/bash
# round 2, second pass: window stem + engine with the padded stem input, ncu captures of the TMA kernel, step breakdown
mkdir -p gpurun_out
LOG=gpurun_out/r2_conv_tma2.log
: > $LOG
COCLR_TMA_DEBUG=1 timeout 600 python -m pytest tests/test_conv_tma_gpu.py -q -p no:cacheprovider 2>&1 | tail -30 >> $LOG
echo "== tma suite rc $?" >> $LOG
timeout 900 python -m pytest tests/test_input_gpu.py tests/test_infonce_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -15 >> $LOG
echo "== engine suites rc $?" >> $LOG
for tma in 1 0; do
  export COCLR_TMA=$tma
  echo "---- COCLR_TMA=$tma" >> $LOG
  timeout 120 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 120 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 120 python tests/tools/run_one_conv.py dgrad 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
done
export COCLR_TMA=1
N="ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 3 -c 1 -f"
STRIDE=2,1,1 timeout 300 $N -o gpurun_out/ncu_r02_tma_stem2_fwd python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 3 2 > /dev/null 2>&1
timeout 300 $N -o gpurun_out/ncu_r02_tma_conv2c1_fwd python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
timeout 300 $N -o gpurun_out/ncu_r02_tma_conv2c1_dgrad python tests/tools/run_one_conv.py dgrad 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
timeout 300 $N -o gpurun_out/ncu_r02_tma_s2d python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 3 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep >> $LOG
timeout 600 python bench.py --no-cpu-baseline --breakdown --steps 6 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench exit $?" >> $LOG
head -c 600 gpurun_out/r2_bench_a.json >> $LOG
head -75 gpurun_out/r2_bench_a.err >> $LOG
tail -100 $LOG
