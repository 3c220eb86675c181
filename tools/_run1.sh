O=gpurun_out/r2p; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q -k "pyramid or session or c1" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
python bench.py --no-cpu --no-sad --no-e2e 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', b['value'], b['ms_per_step'], {k:(round(v['gbs']), v.get('ms_per_launch', v.get('ms_per_launch_group'))) for k,v in b['stages'].items()}, b['parity']['ok'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:pyr_down" -s 6 -c 4 --csv --log-file $O/pyr_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-sad --no-parity > $O/n1.log 2>&1
grep pyr_down $O/pyr_launches.csv | cut -d, -f5,14- | cut -c1-40,150- | tail -4
