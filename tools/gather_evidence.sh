#!/bin/bash
# Regenerates the single-GPU evidence set under profiles/ (run on a B200 box from the repository root,
# e.g. `gpurun --timeout 3000 -- 'bash tools/gather_evidence.sh r02'`; files land in gpurun_out/<tag>/ and
# are copied to profiles/<tag>_* by hand after a look).  Numbers printed under ncu are never bench values:
# every ncu command below profiles a program that was first run plainly.
TAG=${1:-r02}; O=gpurun_out/$TAG; mkdir -p $O   # second argument "core": stop before the microbench / compat / PCIe legs
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > $O/gpu.txt; nproc >> $O/gpu.txt
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
python bench.py > $O/bench_final.json 2> $O/bench_final.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
python bench.py --width 960 --height 540 --frames 30 --batch 30 --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err
python bench.py --impl reference --width 960 --height 540 --frames 30 --batch 30 --steps 10 --warmup 1 > $O/bench_c1_reference_arm.json 2>> $O/bench_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-sad --no-parity > $O/ncu_launch.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
# the five kernels of one 100-frame batch (warm-up step = 15 matching launches, then one batch)
$NCU -k "regex:dct8x8|pyr_down|hbma_strip" -s 15 -c 5 -o $O/step_kernels \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-sad --no-parity > $O/n2.log 2>&1
$NCU -k regex:hbma_rs -s 2 -c 2 -o $O/rs_R16L2 python tools/sweep_hbma.py --ranges 16 --levels 2 --cpu-budget-gabsdiff 0 --out $O/sw_ncu > $O/n3.log 2>&1
$NCU -k "regex:hbma_rs|hbma_ebma_tile" -s 4 -c 4 -o $O/rs_R64L4 python tools/sweep_hbma.py --ranges 64 --levels 4 --cpu-budget-gabsdiff 0 --out $O/sw_ncu > $O/n4.log 2>&1
python tools/sweep_hbma.py --out $O/sweep_hbma > $O/sweep.log 2>&1
python tools/ab_hbma.py --out $O/ab_hbma_strip_vs_tile.json > $O/ab.log 2>&1
python tools/ncu_summary.py $O/step_kernels.ncu-rep > $O/ncu_full_step_kernels.txt 2>&1
if [ "$2" = "core" ]; then ls $O; exit 0; fi
for tb in 8 16 4; do python tools/microbench.py --tb $tb --out $O/microbench_4k_tb$tb.json > $O/mb$tb.log 2>&1; done
python tools/compat_bench.py > $O/compat.json 2> $O/compat.err
python tools/pcie_probe.py --out $O/pcie_probe_n1.json > $O/pcie.log 2>&1
# summaries: python tools/ncu_summary.py $O/step_kernels.ncu-rep > profiles/${TAG}_ncu_full_step_kernels.txt  (etc.)
# multi-GPU (gpurun --gpus N): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#   --master-port 29521 bench.py --gpus N [--width 3840 --height 2160 --frames 600/N]   and   tools/pcie_probe.py
ls $O
