#!/bin/bash
# Round 2 closing evidence pass on the final build (one GPU): tests, smoke, every bench line, ncu launch list + full capture of the
# headline kernel, memcheck, the B=64 tables.  Everything lands in gpurun_out/r2_closing/ (copied into profiles/ afterwards).
set -u
OUT=gpurun_out/r2_closing
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "1 gpu tests"
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
grep -v "^DEBUG\|^INFO" "$OUT/pytest_gpu.log" | tail -8 | tee -a "$OUT/steps.log"
step "2 smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "rc=$?" | tee -a "$OUT/steps.log"
step "3 bench lines"
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "ref rc=$?" | tee -a "$OUT/steps.log"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench_warp.json" 2> "$OUT/bench_warp.err"; echo "warp rc=$?" | tee -a "$OUT/steps.log"
for wl in blur warp_bwd small ingest; do
  timeout 400 python bench.py --workload $wl > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"; echo "$wl rc=$?" | tee -a "$OUT/steps.log"
done
step "4 ncu launch list of bench.py (final build)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_steps3.csv" \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-side-legs > "$OUT/ncu_launches.log" 2>&1
step "5 ncu --set full of the headline kernel at the headline batch"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:warp_fwd_tma.*int.64.*int.32.*int.72" -s 2 -c 1 -o "$OUT/prof_fwd_B256" \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side-legs > "$OUT/ncu_fwd.log" 2>&1; tail -2 "$OUT/ncu_fwd.log" | tee -a "$OUT/steps.log"
step "6 memcheck"
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > "$OUT/memcheck.txt" 2>&1; tail -3 "$OUT/memcheck.txt" | tee -a "$OUT/steps.log"
step "7 tables at B=64"
timeout 300 python tools/bench_modes.py > "$OUT/modes_B64.txt" 2>&1; cp gpurun_out/modes.json "$OUT/modes_B64.json" 2>/dev/null
timeout 300 python tools/bench_filters.py > "$OUT/filters_remap_B64.txt" 2>&1
timeout 300 python tools/bench_family.py > "$OUT/family_B64.txt" 2>&1
timeout 300 python tools/ab_remap.py 64 > "$OUT/ab_remap_B64.txt" 2>&1
timeout 300 python tools/diag_reflection.py > "$OUT/reflection_per_sample.txt" 2>&1
ls -la "$OUT" | tee -a "$OUT/steps.log"
