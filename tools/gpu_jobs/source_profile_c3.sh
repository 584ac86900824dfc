cd $GRAFT_REPO_ROOT
timeout 900 ncu --section SourceCounters --section LaunchStats --clock-control none --import-source on -k regex:k_wsolve -c 1 -o gpurun_out/r2_c3_final -f python tools/gpu_c3_probe.py 200x1000 > gpurun_out/r2_c3_final.log 2>&1
tail -2 gpurun_out/r2_c3_final.log
