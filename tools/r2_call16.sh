OUT=gpurun_out/r2_call16; mkdir -p $OUT
timeout 300 python tools/diag_reflection.py > $OUT/diag_reflection.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:remap_piped -s 2 -c 1 -o $OUT/prof_remap_refl python tools/ncu_targets.py remap_reflection > $OUT/ncu_remap.log 2>&1
