mkdir -p gpurun_out/r2b; O=gpurun_out/r2b
for x in 0 16 1 2 3 5 -3 -16 17; do echo "x=$x"; timeout 20 tools/tma_probe 1920 1088 1920 2 48 32 $x 7 1; done > $O/tma_probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
tail -3 $O/pytest.log; cat $O/tma_probe.log; tail -c 1500 $O/bench.json; tail -3 $O/bench.err
