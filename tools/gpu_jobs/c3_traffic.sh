cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_wsolve -c 1 --csv --log-file gpurun_out/r2_c3_traffic_final.csv python tools/gpu_c3_probe.py 1000x1000 > /dev/null 2>&1
tail -4 gpurun_out/r2_c3_traffic_final.csv
