OUT=gpurun_out/r2_call11; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py -x -q -m gpu -k "remap or row_pass" 2>&1 | tail -5 > $OUT/pytest.txt
timeout 300 python tools/ab_remap.py 64 > $OUT/ab_remap_B64.txt 2>&1
timeout 300 python tools/ab_blur.py 256 sep_rowpair > $OUT/ab_blur_rowpair_B256.txt 2>&1
KB200_REMAP_PIPED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:remap_piped -s 2 -c 1 -o $OUT/prof_remap_piped python tools/ncu_targets.py remap > $OUT/ncu_remap.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_fwd_tma -s 3 -c 1 -o $OUT/prof_reflection python tools/ncu_targets.py reflection > $OUT/ncu_refl.log 2>&1
KB200_SEP_ROWPAIR=1 KB200_SEP_VWALK=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sepfilter_tiled -s 2 -c 1 -o $OUT/prof_blur_rowpair python tools/ncu_targets.py blur11 > $OUT/ncu_blur.log 2>&1
