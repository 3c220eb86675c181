O=gpurun_out/r2l; mkdir -p $O
timeout 1500 python -m pytest tests/test_decode.py tests/test_host_cpp.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for tb in 8 16 4; do python tools/microbench.py --tb $tb --out $O/microbench_4k_tb$tb.json > $O/mb$tb.log 2>&1; python -c "
import json; d=json.load(open('$O/microbench_4k_tb$tb.json'))['results']; print($tb, {k:(round(v['gbs']),round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})"; done
