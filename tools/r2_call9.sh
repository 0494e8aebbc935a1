OUT=gpurun_out/r2_call9; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py -x -q -m gpu -k "reflection or modes or tiled or warp" 2>&1 | tail -5 > $OUT/pytest.txt
timeout 300 python tools/bench_modes.py > $OUT/modes_B64.txt 2>&1; cp gpurun_out/modes.json $OUT/modes_B64.json 2>/dev/null
