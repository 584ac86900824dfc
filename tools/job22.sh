cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "rc=$?"; tail -c 600 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2>> gpurun_out/r2_bench_final.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','us_per_pod','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks'])
for k in ('c2','deployments','c5_one_gpu','consolidation'):
    if k in d: print(k, d[k].get('ms_per_step', d[k].get('ms')), d[k].get('value'), d[k].get('e2e'))
print('encoder', d.get('encoder'))
print('cpu', d.get('cpu_baseline'))
r = json.loads(open('gpurun_out/r2_bench_final_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
