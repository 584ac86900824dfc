cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2>> gpurun_out/r2_bench_final.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','us_per_pod','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
for k in ('c2','deployments','c5_one_gpu','consolidation'):
    if k in d: print(k, d[k].get('ms_per_step', d[k].get('ms')), d[k].get('value'), (d[k].get('e2e') or {}).get('ms_per_step', (d[k].get('e2e') or {}).get('ms')))
print('encoder', d['encoder']['pods_per_s'], d['encoder']['e2e_with_encode_ms_extrapolated'])
print('cpu', d['cpu_baseline']['value'], d['c2']['cpu_baseline']['value'], d['consolidation']['cpu_baseline']['value'])
r = json.loads(open('gpurun_out/r2_bench_final_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'])
PY
