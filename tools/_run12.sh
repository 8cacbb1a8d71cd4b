set -u
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filtered_walk" 2>&1 | tail -4
# (1) launch list of the default bench command, (2) full captures of the two dominant kernels, (3) fp32 split launch list
python bench.py --steps 5 --warmup 3 --extra-batches 1024 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 5 --warmup 3 --extra-batches 1024 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/prof_tensor.py 10000000 768 cosine 10 bf16 1 > gpurun_out/plain_scan.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_stream -s 2 -c 1 -o gpurun_out/prof_scan_r01c -f python tools/prof_tensor.py 10000000 768 cosine 10 bf16 1 > gpurun_out/ncu_scan.log 2>&1
python tools/prof_tensor.py 10000000 768 > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 -o gpurun_out/prof_gemm_r01c -f python tools/prof_tensor.py 10000000 768 > gpurun_out/ncu_gemm.log 2>&1
python tools/prof_tensor.py 1000000 384 cosine 10 f32 1024 > gpurun_out/plain_split.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_split_r01c.csv python tools/prof_tensor.py 1000000 384 cosine 10 f32 1024 > gpurun_out/ncu_split.log 2>&1
tail -n 2 gpurun_out/plain_bench.log gpurun_out/ncu_scan.log gpurun_out/ncu_gemm.log gpurun_out/plain_split.log | cut -c1-300
