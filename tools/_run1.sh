O=gpurun_out/r2j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -k "hbma or sweep or fuzz" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
python tools/sweep_hbma.py --ranges 8,16,32,64 --levels 1,2,3,4,5 --cpu-budget-gabsdiff 0.3 --out $O/sweep > $O/sweep.log 2>&1; grep -E "^\| (8 \| 1|16 \| 2|32 \| 3|64 \| 4|64 \| 5|32 \| 4)" $O/sweep.md
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hbma_ --csv --log-file $O/launches.csv python tools/sweep_hbma.py --ranges 64 --levels 4 --cpu-budget-gabsdiff 0 --out $O/sw_ncu > $O/n1.log 2>&1
grep hbma_ $O/launches.csv | cut -d, -f5,14- | sort | uniq -c | sort -rn | awk 'NR%4==1' | head -8
