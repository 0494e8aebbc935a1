OUT=gpurun_out/r2_call15; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_family_gpu.py tests/test_torch_ops.py -x -q -m gpu 2>&1 | tail -5 > $OUT/pytest.txt
timeout 300 python tools/bench_modes.py > $OUT/modes_B64.txt 2>&1; cp gpurun_out/modes.json $OUT/modes_B64.json 2>/dev/null
timeout 300 python tools/ab_remap.py 64 > $OUT/ab_remap_B64.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:remap_piped -s 2 -c 1 -o $OUT/prof_remap_refl python tools/ncu_targets.py remap_reflection > $OUT/ncu_remap.log 2>&1
