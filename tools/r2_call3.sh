#!/bin/bash
# Round 2, call 3: full GPU suite (0 gated tests, edge-tile backward), cfg4 again, the 256x256 operating point with the direct-call
# path and the augmentation classes, secondary-mode table + ncu captures of remap_tiled / filter2d_tiled / ssim_vwalk / grad_tiled.
set -u
OUT=gpurun_out/r2_call3
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a "$OUT/steps.log"; }
step "1 gpu tests"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"
grep -v "^DEBUG\|^INFO" "$OUT/pytest_gpu.log" | tail -12 | tee -a "$OUT/steps.log"
step "2 cfg4 (edge tiles on the fast path) with side legs"
timeout 400 python bench.py --workload warp_bwd > "$OUT/bench_warp_bwd.json" 2> "$OUT/bench_warp_bwd.err"; echo "rc=$?" | tee -a "$OUT/steps.log"
step "3 256x256 operating point + augmentation classes"
timeout 300 python bench.py --workload small > "$OUT/bench_small.json" 2> "$OUT/bench_small.err"; echo "small rc=$?" | tee -a "$OUT/steps.log"
step "4 secondary modes and families at B=64"
timeout 300 python tools/bench_modes.py > "$OUT/modes_B64.txt" 2>&1
cp gpurun_out/modes.json "$OUT/modes_B64.json" 2>/dev/null
timeout 300 python tools/bench_filters.py > "$OUT/filters_remap_B64.txt" 2>&1
timeout 300 python tools/bench_family.py > "$OUT/family_B64.txt" 2>&1
step "5 ncu --set full of the kernels without a capture yet"
cat > /tmp/ncu_targets.py <<'PY'
import sys, torch, kornia_b200 as K
which = sys.argv[1]
B = 16
x = torch.rand(B, 3, 1080, 1920, device="cuda")
if which == "remap":
    ys, xs = torch.meshgrid(torch.arange(1080, dtype=torch.float32, device="cuda"), torch.arange(1920, dtype=torch.float32, device="cuda"), indexing="ij")
    r2 = ((xs - 960) / 960) ** 2 + ((ys - 540) / 540) ** 2
    mx = (960 + (xs - 960) * (1 + 0.02 * r2))[None].expand(B, -1, -1).contiguous(); my = (540 + (ys - 540) * (1 + 0.02 * r2))[None].expand(B, -1, -1).contiguous()
    f = lambda: K.remap(x, mx, my, align_corners=True)
elif which == "filter2d":
    k = torch.randn(1, 7, 7, device="cuda"); f = lambda: K.filter2d(x, k)
elif which == "ssim":
    y = x.flip(-1).contiguous(); f = lambda: K.metrics.ssim(x, y, 11)
elif which == "grad":
    f = lambda: K.filters.spatial_gradient(x, "sobel", 1)
elif which == "bicubic":
    import bench
    M = bench.make_homographies(B, 3).cuda(); f = lambda: K.warp_perspective(x, M, (1080, 1920), mode="bicubic")
for _ in range(3): f()
torch.cuda.synchronize()
PY
for t in remap:remap_tiled filter2d:filter2d_tiled ssim:ssim_vwalk grad:grad_tiled bicubic:warp_fwd_tma; do
  name=${t%%:*}; kern=${t##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 1 -o "$OUT/prof_$name" python /tmp/ncu_targets.py $name > "$OUT/ncu_$name.log" 2>&1
done
ls -la "$OUT" | tee -a "$OUT/steps.log"
