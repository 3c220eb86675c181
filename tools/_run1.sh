O=gpurun_out/r2q; mkdir -p $O
timeout 300 python bench.py --width 960 --height 540 --frames 30 --batch 30 --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "c1 rc=$?"
timeout 300 python bench.py --impl reference --width 960 --height 540 --frames 30 --batch 30 --steps 10 --warmup 1 > $O/bench_c1_reference_arm.json 2>> $O/bench_c1.err
timeout 300 python bench.py --impl reference --width 960 --height 540 --frames 30 --batch 30 --steps 10 --warmup 1 2>/dev/null | head -c 0
taskset -c 0 python bench.py --impl reference --width 960 --height 540 --frames 30 --batch 30 --steps 3 --warmup 1 > $O/bench_c1_reference_arm_1core.json 2>> $O/bench_c1.err
taskset -c 0 python bench.py --impl reference --steps 1 --warmup 0 --frames 31 > $O/bench_c2_reference_arm_1core.json 2>> $O/bench_c1.err
python - <<'P'
import json
O='gpurun_out/r2q/'
for f in ('bench_c1.json','bench_c1_reference_arm.json','bench_c1_reference_arm_1core.json','bench_c2_reference_arm_1core.json'):
    b=json.loads(open(O+f).read().strip().splitlines()[-1])
    print(f, b['config']['workload'][:45], 'value', round(b['value'],1), 'e2e', b['e2e'] and round(b['e2e']['value'],1), 'cores', b['cpu_baseline'] and b['cpu_baseline']['cores'], 'parity', b.get('parity') and b['parity']['ok'])
P
