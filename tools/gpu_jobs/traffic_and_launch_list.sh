cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_slot_kats.py -m gpu -q 2>&1 | tail -4
echo "== MIN_CTAS=2 (default)"; python tools/gpu_c4_probe.py 3 | tail -2
echo "== MIN_CTAS=3"; KP_LIB_PATH=$PWD/build/libkarpsolve_c3.so python tools/gpu_c4_probe.py 3 | tail -2
echo "== MIN_CTAS=4"; KP_LIB_PATH=$PWD/build/libkarpsolve_c4.so python tools/gpu_c4_probe.py 3 | tail -2
# DRAM traffic of the two dominant kernels (evidence for roofline.traffic), one launch each
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_consolidate -c 1 --csv --log-file gpurun_out/r2_c4_traffic.csv python tools/gpu_c4_probe.py 1 > /dev/null 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_wsolve -c 1 --csv --log-file gpurun_out/r2_c3_traffic.csv python tools/gpu_c3_probe.py 1000x1000 > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_wsolve -c 1 --csv --log-file gpurun_out/r2_c2_traffic.csv python tools/gpu_c2_probe.py > /dev/null 2>&1
tail -4 gpurun_out/r2_c4_traffic.csv gpurun_out/r2_c3_traffic.csv gpurun_out/r2_c2_traffic.csv
# launch list of the bench command (short form), for profiles/
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --apps 100 --no-c5 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
wc -l gpurun_out/r2_launches.csv
