#!/bin/bash
set -u
OUT=gpurun_out/r2_call7
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "1 gpu tests (backward variants, torch ops incl. CUDA graph + half precision)"
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_torch_ops.py tests/test_parity_gpu.py -m gpu -q -rs -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
grep -v "^DEBUG\|^INFO" "$OUT/pytest_gpu.log" | tail -12 | tee -a "$OUT/steps.log"
step "2 cfg4: column-pair lanes vs stride-1 lanes (interleaved, 3 reps)"
for rep in 1 2 3; do for v in 0 1; do
  KB200_BWD_STRIDE1=$v timeout 300 python bench.py --workload warp_bwd --no-side-legs > "$OUT/bench_warp_bwd_s${v}_$rep.json" 2>> "$OUT/bench.err"
  python -c "import json;d=json.load(open('$OUT/bench_warp_bwd_s${v}_$rep.json'));print('stride1=$v rep $rep step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'])" | tee -a "$OUT/steps.log"
done; done
step "3 ncu of the stride-1 variant"
KB200_BWD_STRIDE1=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_bwd_tma2 -s 1 -c 1 -o "$OUT/prof_bwd_s1" \
  python bench.py --workload warp_bwd --batch 32 --steps 1 --warmup 1 --no-side-legs > "$OUT/ncu_bwd.log" 2>&1
step "4 256x256 with CUDA graphs"
timeout 300 python bench.py --workload small > "$OUT/bench_small.json" 2> "$OUT/bench_small.err"; echo "small rc=$?" | tee -a "$OUT/steps.log"
ls -la "$OUT" | tee -a "$OUT/steps.log"
