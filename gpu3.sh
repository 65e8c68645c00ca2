cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dist_run.py -x -q -m gpu > gpurun_out/r2c_tests_dist.log 2>&1; echo "dist tests rc=$?"
tail -5 gpurun_out/r2c_tests_dist.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_dist_run.py > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2c_tests.log
