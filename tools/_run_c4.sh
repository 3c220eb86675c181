# usage: bash tools/_run_c4.sh N FRAMES_PER_GPU
N=$1; F=$2; O=gpurun_out/r2n; mkdir -p $O
free -g | head -2 | tail -1; nproc
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --width 3840 --height 2160 --frames $F --steps 5 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err; echo "C4 N=$N rc=$?"; tail -2 $O/bench_c4_n$N.err | cut -c1-300
python - <<P
import json
b=json.loads(open('$O/bench_c4_n$N.json').read().strip().splitlines()[-1])
print(b['n_gpus'], b['config']['workload'][:60], 'value', round(b['value']), 'e2e', b['e2e'] and {k:b['e2e'][k] for k in ('value','frac') if k in b['e2e']}, 'parity', b['parity'] and (b['parity']['ok'], b['parity']['mv_frames'], b['parity']['dct_max_abs_err']))
P
