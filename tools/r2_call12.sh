OUT=gpurun_out/r2_call12; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_family_gpu.py -x -q -m gpu -k "remap or undistort" 2>&1 | tail -5 > $OUT/pytest.txt
timeout 300 python tools/ab_remap.py 64 > $OUT/ab_remap_B64.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file $OUT/launches_reflection.csv python tools/ncu_targets.py reflection > $OUT/ncu_refl_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file $OUT/launches_zeros.csv python tools/ncu_targets.py zeros > $OUT/ncu_zeros_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:remap_piped -s 2 -c 1 -o $OUT/prof_remap_piped python tools/ncu_targets.py remap > $OUT/ncu_remap.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:warp_fwd_tma.*int.64.*int.32.*int.72" -s 1 -c 1 -o $OUT/prof_zeros python tools/ncu_targets.py zeros > $OUT/ncu_zeros.log 2>&1
