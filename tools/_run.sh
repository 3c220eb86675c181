python tools/sweep_hbma.py --out gpurun_out/sweep_hbma_v7 > gpurun_out/sweep_v7.log 2>&1
cat gpurun_out/sweep_hbma_v7.md
ncu --set full --import-source on --clock-control none -k regex:hbma_ebma_tile -c 1 -f -o gpurun_out/prof_tile_r32 python tools/sweep_hbma.py --levels 1 --ranges 32 --cpu-budget-gabsdiff 0 --frames 5 --out gpurun_out/tmp_sw > gpurun_out/ncu_tile_r32.log 2>&1
