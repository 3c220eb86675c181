mkdir -p gpurun_out/r2c; O=gpurun_out/r2c
ncu --set full --clock-control none --import-source on -k "regex:pyr_down" -s 4 -c 2 -o $O/pyr2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-sad --no-parity > $O/n1.log 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
