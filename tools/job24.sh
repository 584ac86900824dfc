cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --apps 100 --no-c5 --no-cpu-baseline --no-deployments > gpurun_out/r2_launches_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_final.csv', errors='replace')) if len(r) > 10 and r[0].isdigit()]
agg = collections.Counter(); n = collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:60]
    try: v = float(r[-1].replace(',', ''))
    except ValueError: continue
    agg[name] += v; n[name] += 1
tot = sum(agg.values())
for k, v in agg.most_common(8): print(f"{100*v/tot:6.2f}% {n[k]:4d}x {k}")
PY
