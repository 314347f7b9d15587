bash scripts/gpu_round_check.sh 2>&1 | tail -22
timeout 300 python bench.py --eval-sweep --sweep-items 5000000 --sweep-heads 1,12 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('sweep',d['items'],d['heads'],round(d['value']),round(d['ms_per_batch'],2),round(d['useful_tflops']),round(d['table_gbs']))
"
