set -u
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash tools/run_configs.sh > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    c=d['config']; r=d['roofline']
    print(c['workload'][21:70],'sel',c.get('where_selectivity'),'| QPS %.0f ms %.3f e2e %.0f | %s %.0f %.2f'%(d['value'],d['ms_per_step'],d['e2e']['value'],r['bound'],r['achieved'],r['frac']))
PY
