set -u
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "check rc=$?"; grep -v "^W\|^\*" gpurun_out/multi_gpu_check.log | tail -15
for FE in 1 0; do
RAG_B200_FUSED_EXCHANGE=$FE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --rows 2500000 --steps 500 --warmup 20 --extra-batches "" > gpurun_out/bench_2gpu_fe$FE.log 2>&1; echo "bench fe=$FE rc=$?"; tail -1 gpurun_out/bench_2gpu_fe$FE.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['p50_ms'], d['e2e']['value'], d['config']['sharding'], d['gpu_launches'])"
done
