set -u
run() { python bench.py --no-cpu-baseline --extra-batches "" "$@" | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['config']['workload'][21:80], '| QPS %.0f ms %.4f p50 %.4f e2e %.0f | %s %.0f %.2f | launches %d'%(d['value'],d['ms_per_step'],d['p50_ms'],d['e2e']['value'],r['bound'],r['achieved'],r['frac'],d['gpu_launches']))"; }
run --rows 10000000 --dim 768 --steps 300 --warmup 20 --verify
RAG_B200_PDL=0 run --rows 10000000 --dim 768 --steps 300 --warmup 20
run --rows 1250000 --dim 768 --steps 1000 --warmup 20 --verify
RAG_B200_PDL=0 run --rows 1250000 --dim 768 --steps 1000 --warmup 20
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
