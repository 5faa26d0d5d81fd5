set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.txt
tail -12 gpurun_out/r02k_pytest.txt
timeout 900 python tools/fuzz_pruned_vs_brute.py --car-steps 2e9 --steps 4 --throws 12 --out gpurun_out/r02k_fuzz.json > gpurun_out/r02k_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02k_fuzz.log
tail -3 gpurun_out/r02k_fuzz.log | cut -c1-900
timeout 600 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" > gpurun_out/r02k_smoke.txt 2>&1; tail -2 gpurun_out/r02k_smoke.txt
