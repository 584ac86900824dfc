cd $GRAFT_REPO_ROOT
timeout 600 ncu --section SourceCounters --section LaunchStats --clock-control none --import-source on -k regex:k_wsolve -c 1 -o gpurun_out/r2_c2_v3 -f python tools/gpu_c2_probe.py > gpurun_out/r2_c2_v3.log 2>&1
tail -3 gpurun_out/r2_c2_v3.log
