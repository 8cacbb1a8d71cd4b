set -u
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.log
bash tools/run_configs.sh > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    c=d['config']; r=d['roofline']
    print(c['workload'][21:78],'sel',c.get('where_selectivity'),'| QPS %.0f ms %.4f p50 %.4f e2e %.0f | %s %.0f %s %.2f'%(d['value'],d['ms_per_step'],d['p50_ms'],d['e2e']['value'],r['bound'],r['achieved'],r['unit'],r['frac']))
PY
