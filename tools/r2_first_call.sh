#!/bin/bash
# First gpurun call of round 2 (one box, one GPU; budget ~30-35 min of box time):
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r2_first_call.sh'
#
# Round 1 ended with the GPU budget spent before (a) the wider-caller tests, (b) the kernels of DESIGN.md section 9
# and (c) the ncu launch list of the final build could run on hardware.  This script collects all of it in one call
# and leaves everything under gpurun_out/r2_first/ (copy what is to be judged into profiles/).
# Every step is bounded by its own timeout and failures do not stop the following steps.
# Order: everything about the DEFAULT build first (tests, bench lines, ncu launch list and captures), then the device code
# that has never met hardware -- so a kernel that hangs or faults there cannot cost the evidence of the verified paths.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }

step "1 gpu tests of the default build"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
tail -3 "$OUT/pytest_gpu.log" | tee -a "$OUT/steps.log"

step "2 bench lines (headline, blur, fwd+bwd) and the CPU arm"
for wl in warp blur warp_bwd; do
  timeout 400 python bench.py --workload $wl > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"; echo "$wl rc=$?" | tee -a "$OUT/steps.log"
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"

step "3 ncu launch list of bench.py on this build (shares, not absolutes)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_steps3.csv" \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1

step "4 ncu --set full of the three BASELINE kernels (small batches: ~40 replays per launch)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_bwd_tma -s 1 -c 1 -o "$OUT/prof_bwd" \
  python bench.py --workload warp_bwd --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/ncu_bwd.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sepfilter_tiled -s 1 -c 1 -o "$OUT/prof_blur" \
  python bench.py --workload blur --batch 16 --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/ncu_blur.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_fwd_tma -s 2 -c 1 -o "$OUT/prof_fwd" \
  python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/ncu_fwd.log" 2>&1

step "5 unverified kernels (section 9): parity against the verified paths"
KB200_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/test_unverified_gpu.py tests/test_ingest_gpu.py -m gpu -q -p no:cacheprovider > "$OUT/pytest_unverified.log" 2>&1
echo "rc=$?" >> "$OUT/pytest_unverified.log"
tail -15 "$OUT/pytest_unverified.log" | tee -a "$OUT/steps.log"

step "6 the whole GPU suite THROUGH the opt-in kernels (every golden / fp64 / gradcheck / full-size test of round 1)"
KB200_OPTIN=all timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu_optin.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu_optin.log"
tail -5 "$OUT/pytest_gpu_optin.log" | tee -a "$OUT/steps.log"

step "7 timings of the unverified variants against the defaults"
timeout 400 python tools/bench_unverified.py > "$OUT/bench_unverified.txt" 2>&1; echo "rc=$?" | tee -a "$OUT/steps.log"
timeout 300 python tools/bench_family.py > "$OUT/family_B64.txt" 2>&1
timeout 300 python tools/bench_vs_torch.py > "$OUT/vs_torch.txt" 2>&1   # the BASELINE configs against torch eager on the same GPU
# the two BASELINE workloads with their opt-in kernels: full bench lines (roofline leg included)
KB200_SEP_VWALK=1 timeout 400 python bench.py --workload blur --no-cpu-baseline > "$OUT/bench_blur_vwalk.json" 2> "$OUT/bench_blur_vwalk.err"
KB200_BWD_V2=1 timeout 400 python bench.py --workload warp_bwd --no-cpu-baseline > "$OUT/bench_warp_bwd_v2.json" 2> "$OUT/bench_warp_bwd_v2.err"
timeout 400 python bench.py --workload ingest --no-cpu-baseline > "$OUT/bench_ingest.json" 2> "$OUT/bench_ingest.err"; echo "ingest rc=$?" | tee -a "$OUT/steps.log"

step "8 compute-sanitizer memcheck: default kernels + ingest warps, then the same workload through the opt-in kernels"
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > "$OUT/memcheck_default.txt" 2>&1; tail -3 "$OUT/memcheck_default.txt" | tee -a "$OUT/steps.log"
KB200_OPTIN=all timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > "$OUT/memcheck_optin.txt" 2>&1; tail -3 "$OUT/memcheck_optin.txt" | tee -a "$OUT/steps.log"
ls -la "$OUT" | tee -a "$OUT/steps.log"
