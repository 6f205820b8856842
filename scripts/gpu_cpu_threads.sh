#!/bin/bash
# how does the CPU (oracle) arm scale with threads on the GPU box's host?
python -c "import torch, os; print('default threads', torch.get_num_threads(), 'cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))"
cat /sys/fs/cgroup/cpu.max 2>/dev/null
for t in 0 8 16 32 64 128; do COCLR_CPU_THREADS=$t timeout 300 python bench.py --impl reference --steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads', $t, 'clips/s', round(d['value'],3), 'ms', round(d['ms_per_step'],1), 'cores', d['cpu_baseline']['cores'])"; done
