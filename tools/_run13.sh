set -u
run() { python bench.py --no-cpu-baseline --extra-batches "" "$@" | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['config']; print(c['workload'][21:80], 'sel',c.get('where_selectivity'), '| QPS %.0f ms %.4f p50 %.4f e2e %.0f | %s %.0f %.2f | clk %s'%(d['value'],d['ms_per_step'],d['p50_ms'],d['e2e']['value'],r['bound'],r['achieved'],r['frac'],d['clocks'].get('sm_mhz')))"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run --batch 1024 --steps 20 --warmup 3 --verify
run --rows 25000000 --dim 384 --space l2 --batch 1024 --steps 10 --warmup 3
run --rows 25000000 --dim 384 --space l2 --batch 1024 --k 100 --steps 10 --warmup 3
run --rows 1000000 --dim 384 --dtype f32 --batch 1024 --steps 50 --warmup 5
run --rows 1000000 --dim 384 --dtype f32 --batch 32 --steps 50 --warmup 5
run --batch 32 --steps 50 --warmup 3
run --batch 128 --steps 50 --warmup 3
