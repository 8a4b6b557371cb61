set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_chain_many.py -x -q 2>&1 | tail -15 > gpurun_out/r2c_new_tests.log; tail -15 gpurun_out/r2c_new_tests.log
( SHIPENV_NO_CHAIN=1 timeout 300 python tools/chain_bench.py
  timeout 300 python tools/chain_bench.py ) > gpurun_out/r2c_bench.log 2>&1
cat gpurun_out/r2c_bench.log
