set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tools/micro/f32x2_rate > gpurun_out/r02b_f32x2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:race_rollout_fused -s 8 -c 1 -o gpurun_out/prof_r02b_fused -f python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02b_ncu.log 2>&1
tail -15 gpurun_out/r02b_pytest.txt
cat gpurun_out/r02b_f32x2.txt
