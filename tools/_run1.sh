O=gpurun_out/r2g; mkdir -p $O
timeout 1500 python -m pytest tests/test_baseline_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "sweep or session or c1" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
python bench.py --no-cpu --no-sad --steps 5 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', b['value'], 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'], b['parity']['ok'])"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:hbma_rs -s 2 -c 2 -o $O/rs_R16L2 python tools/sweep_hbma.py --ranges 16 --levels 2 --cpu-budget-gabsdiff 0 --out $O/sw_ncu > $O/n1.log 2>&1
$NCU -k regex:hbma_rs -s 4 -c 4 -o $O/rs_R64L4 python tools/sweep_hbma.py --ranges 64 --levels 4 --cpu-budget-gabsdiff 0 --out $O/sw_ncu > $O/n2.log 2>&1
python tools/sweep_hbma.py --ranges 32,64 --levels 1,2 --out $O/sweep_gold > $O/sweep.log 2>&1; tail -5 $O/sweep_gold.md
