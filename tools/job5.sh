cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
tail -c 6000 gpurun_out/bench_r2_a.json; tail -5 gpurun_out/bench_r2_a.err
