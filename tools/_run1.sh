O=gpurun_out/r2o; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -k "session or encoder or app or c1 or c4 or stream or shard" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
python bench.py --no-cpu --no-sad --steps 5 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', b['value'], 'e2e', {k:b['e2e'][k] for k in ('value','frac','ms_per_step')}, b['e2e']['copy_only_ceiling']['value'], b['parity']['ok'])"
