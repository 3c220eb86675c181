O=gpurun_out/r2n; mkdir -p $O
free -g | head -2 | tail -1; nproc
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29531 tools/pcie_probe.py --out $O/pcie_probe_n8.json > $O/pcie8.log 2>&1; python -c "
import json; d=json.load(open('$O/pcie_probe_n8.json')); print({m:(round(v['total_gbs_aggregate'],1), round(v['frames_per_s_equivalent'])) for m,v in d['modes'].items()})"
timeout 600 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c2_n8.json 2> $O/bench_c2_n8.err; echo "C2 N=8 rc=$?"
python - <<P
import json
b=json.loads(open('$O/bench_c2_n8.json').read().strip().splitlines()[-1])
print(b['n_gpus'], 'value', round(b['value']), 'e2e', {k:b['e2e'][k] for k in ('value','frac','ms_per_step') if k in b['e2e']}, b['e2e'].get('copy_only_ceiling'), 'parity', (b['parity']['ok'], b['parity']['mv_frames']))
P
bash tools/_run_c4.sh 8 75 2>&1 | tail -2
# sharded application: 4K subsequence, 1 device vs 4 devices, byte identity
python - <<'P'
import sys, os, subprocess, time
sys.path.insert(0, 'scalable-video-codec_b200')
from svc_b200.synth import SyntheticSequence
w, h, n = 3840, 2160, 13
fr = SyntheticSequence(w, h, n, seed=1234).frames()
open('/dev/shm/in4k.bgr', 'wb').write(fr.tobytes())
enc = 'scalable-video-codec_b200/bin/svc_encoder'
for dev, out in (('0', '/dev/shm/o1.svc'), ('0,1,2,3', '/dev/shm/o4.svc'), ('0,1,2,3,4,5,6,7', '/dev/shm/o8.svc')):
    t0 = time.time()
    r = subprocess.run([enc, '--width', str(w), '--height', str(h), '--frames', str(n), '--devices', dev, '--out', out, '--verbose', '0', '--seed', '7', '/dev/shm/in4k.bgr'], capture_output=True, text=True)
    print(dev, 'rc', r.returncode, round(time.time() - t0, 2), 's', r.stderr[-200:])
a = open('/dev/shm/o1.svc', 'rb').read()
print('bytes', len(a), 'identical 1 vs 4:', a == open('/dev/shm/o4.svc', 'rb').read(), '1 vs 8:', a == open('/dev/shm/o8.svc', 'rb').read())
P
