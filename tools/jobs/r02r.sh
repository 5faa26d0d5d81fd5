set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.txt
tail -5 gpurun_out/r02r_pytest.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02r_bench20_full.json 2> gpurun_out/r02r_bench20_full.err; echo "rc=$?" >> gpurun_out/r02r_bench20_full.err
tail -5 gpurun_out/r02r_bench20_full.err
python tools/show_bench.py gpurun_out/r02r_bench20_full.json
