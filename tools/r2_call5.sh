#!/bin/bash
# Round 2, call 5: GPU suite after the align_corners=False corrections, the default bench (overlapped e2e, uint8-frames e2e), the
# reference arm, cfg3 / cfg4 lines.
set -u
OUT=gpurun_out/r2_call5
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "1 gpu tests"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
grep -v "^DEBUG\|^INFO" "$OUT/pytest_gpu.log" | tail -25 | tee -a "$OUT/steps.log"
step "2 smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "rc=$?" | tee -a "$OUT/steps.log"; tail -6 "$OUT/smoke.log" | tee -a "$OUT/steps.log"
step "3 bench (default) and reference arm"
timeout 600 python bench.py > "$OUT/bench_warp.json" 2> "$OUT/bench_warp.err"; echo "warp rc=$?" | tee -a "$OUT/steps.log"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "ref rc=$?" | tee -a "$OUT/steps.log"
step "4 cfg3 / cfg4"
timeout 400 python bench.py --workload blur > "$OUT/bench_blur.json" 2> "$OUT/bench_blur.err"; echo "blur rc=$?" | tee -a "$OUT/steps.log"
timeout 400 python bench.py --workload warp_bwd > "$OUT/bench_warp_bwd.json" 2> "$OUT/bench_warp_bwd.err"; echo "bwd rc=$?" | tee -a "$OUT/steps.log"
ls -la "$OUT" | tee -a "$OUT/steps.log"
