set -u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check8.log 2>&1; echo "check rc=$?"; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/multi_gpu_check8.log | grep -v "^$" | tail -6
sum() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['n_gpus'],'gpus', d['config']['workload'][21:], '| QPS %.0f ms %.4f p50 %.4f e2e %.0f (p50 %.4f) | launches %d | %s %.0f %.2f | %s'%(d['value'],d['ms_per_step'],d['p50_ms'],d['e2e']['value'],d['e2e']['p50_ms'],d['gpu_launches'],r['bound'],r['achieved'],r['frac'],d['config']['sharding'][:50])); print('   regimes', [(r['batch'], round(r['value']), round(r['roofline']['frac'],2)) for r in d.get('regimes',[])], d['clocks'])"; }
timeout 600 $TR --master-port 29512 bench.py --gpus 8 > gpurun_out/bench_8gpu_b1.log 2>&1; echo "bench rc=$?"; sum gpurun_out/bench_8gpu_b1.log
timeout 900 $TR --master-port 29514 bench.py --gpus 8 --rows 200000000 --dim 384 --space l2 --batch 1024 --k 100 --steps 10 --warmup 3 --extra-batches "" > gpurun_out/bench_8gpu_config5.log 2>&1; echo "bench rc=$?"; sum gpurun_out/bench_8gpu_config5.log
