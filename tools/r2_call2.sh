#!/bin/bash
# Round 2, call 2: the whole GPU suite on the promoted defaults (0 gated tests), the three tiled backward kernels side by side
# (cfg4), the new bench legs (parity table, torch eager / torch.compile on the same GPU, 256x256 operating point), ncu of V3.
set -u
OUT=gpurun_out/r2_call2
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "1 gpu tests"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
grep -v "^DEBUG\|^INFO" "$OUT/pytest_gpu.log" | tail -12 | tee -a "$OUT/steps.log"
step "2 cfg4 with warp_bwd_tma2 / tma3 (4-pixel units) / tma3 (2-pixel units)"
for v in 0 1 2; do
  KB200_BWD_V3=$v timeout 300 python bench.py --workload warp_bwd --no-side-legs > "$OUT/bench_warp_bwd_v3_$v.json" 2> "$OUT/bench_warp_bwd_v3_$v.err"; echo "v3=$v rc=$?" | tee -a "$OUT/steps.log"
done
step "3 ncu of warp_bwd_tma3"
KB200_BWD_V3=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_bwd_tma3 -s 1 -c 1 -o "$OUT/prof_bwd3" \
  python bench.py --workload warp_bwd --batch 32 --steps 1 --warmup 1 --no-side-legs > "$OUT/ncu_bwd3.log" 2>&1
step "4 headline with side legs, small operating point, blur"
timeout 600 python bench.py > "$OUT/bench_warp.json" 2> "$OUT/bench_warp.err"; echo "warp rc=$?" | tee -a "$OUT/steps.log"
timeout 300 python bench.py --workload small > "$OUT/bench_small.json" 2> "$OUT/bench_small.err"; echo "small rc=$?" | tee -a "$OUT/steps.log"
timeout 400 python bench.py --workload blur > "$OUT/bench_blur.json" 2> "$OUT/bench_blur.err"; echo "blur rc=$?" | tee -a "$OUT/steps.log"
ls -la "$OUT" | tee -a "$OUT/steps.log"
