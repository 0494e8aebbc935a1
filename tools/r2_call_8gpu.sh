#!/bin/bash
# Round 2, 8-GPU call (charged 8x): host topology, the headline at N=8 with NUMA-bound pinned buffers (e2e scaling), the
# scatter/gather form of cfg5 over NCCL, and the e2e pipeline with unbound buffers for contrast.
set -u
OUT=gpurun_out/r2_8gpu
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "0 topology"
( lscpu | head -25; echo; cat /sys/devices/system/node/node*/cpulist; echo; nvidia-smi topo -m; echo; free -g; for n in /sys/devices/system/node/node*; do echo $n; head -4 $n/meminfo; done ) > "$OUT/topology.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step "1 headline N=8 (NUMA-bound e2e)"
timeout 500 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > "$OUT/bench_n8.json" 2> "$OUT/bench_n8.err"; echo "rc=$?" | tee -a "$OUT/steps.log"
step "2 headline N=8, binding disabled (round-1 behaviour)"
KB200_NO_NUMA_BIND=1 timeout 500 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 > "$OUT/bench_n8_unbound.json" 2> "$OUT/bench_n8_unbound.err"; echo "rc=$?" | tee -a "$OUT/steps.log"
step "3 cfg5 scatter/gather over NCCL, global B = 512 on rank 0"
timeout 500 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --workload scatter_gather --global-batch 512 --steps 5 --warmup 2 > "$OUT/bench_sg_n8.json" 2> "$OUT/bench_sg_n8.err"; echo "rc=$?" | tee -a "$OUT/steps.log"
step "4 N=4 headline (one socket)"
timeout 400 $TR --nproc-per-node 4 --master-port 29614 bench.py --gpus 4 --steps 10 --warmup 3 > "$OUT/bench_n4.json" 2> "$OUT/bench_n4.err"; echo "rc=$?" | tee -a "$OUT/steps.log"
tail -c 600 "$OUT"/*.err | tee -a "$OUT/steps.log"
ls -la "$OUT" | tee -a "$OUT/steps.log"
