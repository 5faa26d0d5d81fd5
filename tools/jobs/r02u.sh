set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_rollout_gpu.py tests/test_race_gpu.py -q -x -k "fixture or host_rollout or validity or other_track_lengths" > gpurun_out/r02u_plain.txt 2>&1 &&
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fused_rollout_gpu.py tests/test_race_gpu.py -q -x -k "fixture or host_rollout or validity or other_track_lengths" > gpurun_out/r02u_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02u_memcheck.txt
tail -3 gpurun_out/r02u_plain.txt; tail -12 gpurun_out/r02u_memcheck.txt
