set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_race_gpu.py tests/test_full_size_gpu.py tests/test_helpers_gpu.py -q -x > gpurun_out/r02s_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest.txt
tail -8 gpurun_out/r02s_pytest.txt
timeout 300 python tools/reset_perf.py > gpurun_out/r02s_reset.txt 2>&1; tail -12 gpurun_out/r02s_reset.txt
