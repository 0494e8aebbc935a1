#!/bin/bash
# Round 2, call 6: GPU suite (reflection / fill on the forward fast path, half-precision pass-through), the mode table, an ncu
# capture of the reflection instantiation.
set -u
OUT=gpurun_out/r2_call6
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "1 gpu tests"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
grep -v "^DEBUG\|^INFO" "$OUT/pytest_gpu.log" | tail -15 | tee -a "$OUT/steps.log"
step "2 modes at B=64"
timeout 300 python tools/bench_modes.py > "$OUT/modes_B64.txt" 2>&1; cp gpurun_out/modes.json "$OUT/modes_B64.json" 2>/dev/null; cat "$OUT/modes_B64.txt" | tee -a "$OUT/steps.log"
step "3 ncu reflection / fill"
for name in reflection fill; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:warp_fwd_tma -s 3 -c 1 -o "$OUT/prof_$name" python tools/ncu_targets.py $name > "$OUT/ncu_$name.log" 2>&1
  tail -1 "$OUT/ncu_$name.log" | tee -a "$OUT/steps.log"
done
step "4 headline (regression check of the shared kernel source)"
timeout 400 python bench.py --no-side-legs --no-cpu-baseline > "$OUT/bench_warp.json" 2> "$OUT/bench_warp.err"; echo "rc=$?" | tee -a "$OUT/steps.log"
ls -la "$OUT" | tee -a "$OUT/steps.log"
