#!/bin/bash
# Round 2, call 4: GPU suite (new same-device parity tests), ncu --set full of the kernels without a capture yet, the two one-pass
# blur kernels A/B in one process at the cfg3 batch, the 256x256 operating point with device-side times.
set -u
OUT=gpurun_out/r2_call4
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "1 gpu tests"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
grep -v "^DEBUG\|^INFO" "$OUT/pytest_gpu.log" | tail -25 | tee -a "$OUT/steps.log"
step "2 blur A/B"
timeout 300 python tools/ab_blur.py 256 > "$OUT/ab_blur_B256.txt" 2>&1; tail -6 "$OUT/ab_blur_B256.txt" | tee -a "$OUT/steps.log"
step "3 256x256 operating point"
timeout 300 python bench.py --workload small > "$OUT/bench_small.json" 2> "$OUT/bench_small.err"; echo "small rc=$?" | tee -a "$OUT/steps.log"
step "4 ncu --set full"
for t in remap:remap_tiled filter2d:filter2d_tiled ssim:ssim_vwalk grad:grad_tiled bicubic:warp_fwd_tma reflection:warp_fwd_tma ingest:warp_u8_tiled; do
  name=${t%%:*}; kern=${t##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 1 -o "$OUT/prof_$name" python tools/ncu_targets.py $name > "$OUT/ncu_$name.log" 2>&1
  tail -1 "$OUT/ncu_$name.log" | tee -a "$OUT/steps.log"
done
step "5 ncu launch list of the default bench on this build"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_steps3.csv" \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-side-legs > "$OUT/ncu_launches.log" 2>&1
ls -la "$OUT" | tee -a "$OUT/steps.log"
