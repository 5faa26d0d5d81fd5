set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export GLG_LIB_PATH=$GRAFT_REPO_ROOT/tools/ab/libglg_s1.so
timeout 900 python -m pytest tests/test_fused_rollout_gpu.py tests/test_race_gpu.py tests/test_full_size_gpu.py -q -x > gpurun_out/r02n_s1_pytest_all.txt 2>&1; echo "rc=$?" >> gpurun_out/r02n_s1_pytest_all.txt
tail -3 gpurun_out/r02n_s1_pytest_all.txt
timeout 600 python tools/fuzz_pruned_vs_brute.py --car-steps 3e8 --steps 4 --throws 12 --out gpurun_out/r02n_s1_fuzz.json > gpurun_out/r02n_s1_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02n_s1_fuzz.log
tail -2 gpurun_out/r02n_s1_fuzz.log | cut -c1-400
unset GLG_LIB_PATH
bash tools/jobs/ab.sh r02n s0 s1
