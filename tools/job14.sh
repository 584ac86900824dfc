cd $GRAFT_REPO_ROOT
for v in head v1 v2 v3; do
  echo "== $v"
  KP_LIB_PATH=$PWD/build/libkp_$v.so timeout 300 python tools/gpu_c3_probe.py 300x1000 2>&1 | tail -1
done
