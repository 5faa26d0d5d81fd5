set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest.txt
tail -4 gpurun_out/r02t_pytest.txt
timeout 600 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" > gpurun_out/r02t_smoke.txt 2>&1; tail -2 gpurun_out/r02t_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02t_bench20_full.json 2> gpurun_out/r02t_bench20_full.err; echo "rc=$?" >> gpurun_out/r02t_bench20_full.err
tail -3 gpurun_out/r02t_bench20_full.err
python tools/show_bench.py gpurun_out/r02t_bench20_full.json | cut -c1-600
