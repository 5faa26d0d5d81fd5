set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python tools/fuzz_pruned_vs_brute.py --car-steps 1.2e8 --out gpurun_out/r02h_fuzz.json > gpurun_out/r02h_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02h_fuzz.log
tail -8 gpurun_out/r02h_fuzz.log | cut -c1-1500
