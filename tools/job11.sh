cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err
tail -c 1500 gpurun_out/bench_r2_b.json; tail -3 gpurun_out/bench_r2_b.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.json 2>> gpurun_out/bench_r2_b.err
tail -c 600 gpurun_out/bench_r2_ref.json
