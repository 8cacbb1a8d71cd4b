set -u
python tools/prof_tensor.py 6000000 384 l2 > gpurun_out/plain_b.log 2>&1 && python tools/prof_tensor.py 4000000 768 > gpurun_out/plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 -o gpurun_out/prof_gemm_pair384 -f python tools/prof_tensor.py 6000000 384 l2 > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 -o gpurun_out/prof_gemm_pair768 -f python tools/prof_tensor.py 4000000 768 > gpurun_out/ncu_a.log 2>&1
tail -n 2 gpurun_out/plain_a.log gpurun_out/plain_b.log gpurun_out/ncu_a.log gpurun_out/ncu_c.log
