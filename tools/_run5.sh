set -u
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor" 2>&1 | tail -5
for M in 2 1; do
RAG_B200_TENSOR_MODE=$M timeout 200 python bench.py --no-cpu-baseline --extra-batches "" --batch 1024 --steps 20 --warmup 3 --verify | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('mode $M', d['config']['workload'][21:80], '| QPS %.0f ms %.4f | %s %.0f %.2f'%(d['value'],d['ms_per_step'],r['bound'],r['achieved'],r['frac']), d['clocks'])"
RAG_B200_TENSOR_MODE=$M timeout 200 python bench.py --no-cpu-baseline --extra-batches "" --rows 25000000 --dim 384 --space l2 --batch 1024 --steps 10 --warmup 3 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('mode $M', d['config']['workload'][21:80], '| QPS %.0f ms %.4f | %s %.0f %.2f'%(d['value'],d['ms_per_step'],r['bound'],r['achieved'],r['frac']), d['clocks'])"
done
