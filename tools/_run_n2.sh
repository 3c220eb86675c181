O=gpurun_out/r2m; mkdir -p $O
free -g | head -2; nproc
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29511 tools/pcie_probe.py --out $O/pcie_probe_n2.json > $O/pcie2.log 2>&1; tail -1 $O/pcie2.log | cut -c1-600
$TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"; tail -2 $O/bench_n2.err
python - <<'P'
import json
b=json.loads(open('gpurun_out/r2m/bench_n2.json').read().strip().splitlines()[-1])
print(b['n_gpus'], b['value'], b['e2e'], b['parity'])
P
