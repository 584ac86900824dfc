set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
KP_DEBUG=1 python tools/gpu_c3_probe.py 100x1000 300x1000 1000x1000 > gpurun_out/c3_probe.log 2>&1
tail -20 gpurun_out/c3_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wsolve -c 1 -o gpurun_out/r2_c3_base -f python tools/gpu_c3_probe.py 200x1000 > gpurun_out/r2_c3_base.log 2>&1
tail -5 gpurun_out/r2_c3_base.log
