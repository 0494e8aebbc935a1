#!/bin/bash
# Final-build check (one GPU): the whole GPU suite, smoke, the driver's bench lines.
OUT=gpurun_out/r2_final; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -rs -p no:cacheprovider 2>&1 | tail -4 > $OUT/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_warp.json 2> $OUT/bench_warp.err; echo "bench rc=$?" >> $OUT/pytest.txt
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?" >> $OUT/pytest.txt
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py 2>&1 | tail -3 >> $OUT/pytest.txt
