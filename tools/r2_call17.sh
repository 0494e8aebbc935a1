OUT=gpurun_out/r2_call17; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py -x -q -m gpu 2>&1 | tail -4 > $OUT/pytest.txt
timeout 300 python tools/bench_modes.py > $OUT/modes_B64.txt 2>&1; cp gpurun_out/modes.json $OUT/modes_B64.json 2>/dev/null
timeout 300 python tools/diag_reflection.py > $OUT/diag_reflection.txt 2>&1
