OUT=gpurun_out/r2_call10; mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "remap" 2>&1 | tail -5 > $OUT/pytest.txt
timeout 300 python tools/ab_remap.py 64 > $OUT/ab_remap_B64.txt 2>&1
KB200_REMAP_PIPED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:remap_piped -s 2 -c 1 -o $OUT/prof_remap_piped python tools/ncu_targets.py remap > $OUT/ncu_remap.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_fwd_tma -s 2 -c 1 -o $OUT/prof_reflection python tools/ncu_targets.py reflection > $OUT/ncu_refl.log 2>&1
