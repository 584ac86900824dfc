cd $GRAFT_REPO_ROOT
export KP_FUZZ_SEEDS=12
timeout 1500 compute-sanitizer --tool initcheck --log-file gpurun_out/r2_initcheck.log --print-limit 20 \
  python -m pytest -m gpu -q -x tests/test_fuzz_parity.py tests/test_host_ports.py tests/test_volume_alternatives.py \
  tests/test_truncate_instance_types.py tests/test_consolidation_min_values.py tests/test_reserved_capacity.py \
  tests/test_gpu_slot_kats.py "tests/test_gpu_parity.py::test_deployment_cohorts_parity" "tests/test_gpu_parity.py::test_c1_parity" 2>&1 | tail -5
tail -12 gpurun_out/r2_initcheck.log
