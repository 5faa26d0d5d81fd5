set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02l_bench20_full.json 2> gpurun_out/r02l_bench20_full.err; echo "rc=$?" >> gpurun_out/r02l_bench20_full.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02l_ref.json 2> gpurun_out/r02l_ref.err
timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu --no-extra > gpurun_out/r02l_bench2000.json 2> gpurun_out/r02l_bench2000.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02l_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:race_rollout_fused -s 8 -c 1 -o gpurun_out/prof_r02l_fused -f python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02l_ncu.log 2>&1
python tools/show_bench.py gpurun_out/r02l_bench20_full.json gpurun_out/r02l_ref.json gpurun_out/r02l_bench2000.json
