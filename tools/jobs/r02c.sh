set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.txt
tail -15 gpurun_out/r02c_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02c_bench20.json 2> gpurun_out/r02c_bench20.err
timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu --no-extra > gpurun_out/r02c_bench2000.json 2> gpurun_out/r02c_bench2000.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:race_rollout_fused -s 8 -c 1 -o gpurun_out/prof_r02c_fused -f python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02c_ncu.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02c_bench20_full.json 2> gpurun_out/r02c_bench20_full.err
python tools/show_bench.py gpurun_out/r02c_bench20.json gpurun_out/r02c_bench2000.json gpurun_out/r02c_bench20_full.json
