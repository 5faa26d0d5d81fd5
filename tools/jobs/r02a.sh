set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
tools/micro/f32x2_rate > gpurun_out/r02a_f32x2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench20.json 2> gpurun_out/r02a_bench20.err; echo "rc=$?" >> gpurun_out/r02a_bench20.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02a_bench20_b.json 2> gpurun_out/r02a_bench20_b.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --rollout-mode chained > gpurun_out/r02a_bench20_chained.json 2> gpurun_out/r02a_bench20_chained.err
timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu --no-extra > gpurun_out/r02a_bench2000.json 2> gpurun_out/r02a_bench2000.err
timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu --no-extra --rollout-mode chained > gpurun_out/r02a_bench2000_chained.json 2> gpurun_out/r02a_bench2000_chained.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err
tail -3 gpurun_out/r02a_pytest.txt
cat gpurun_out/r02a_f32x2.txt
