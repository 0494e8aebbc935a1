OUT=gpurun_out/r2_call14; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_reflection64.csv python tools/ncu_targets.py reflection 64 1000 > $OUT/l1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_zeros64.csv python tools/ncu_targets.py zeros 64 1000 > $OUT/l2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:remap_piped -s 2 -c 1 -o $OUT/prof_remap_refl python tools/ncu_targets.py remap_reflection > $OUT/ncu_remap.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:warp_fwd_tma.*int.64.*int.32.*int.72" -s 1 -c 1 -o $OUT/prof_refl64 python tools/ncu_targets.py reflection 64 1000 > $OUT/ncu_refl.log 2>&1
