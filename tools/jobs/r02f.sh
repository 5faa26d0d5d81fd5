set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_helpers_gpu.py -q > gpurun_out/r02f_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.txt
tail -5 gpurun_out/r02f_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02f_bench20.json 2> gpurun_out/r02f_bench20.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --replicas 1 --repeats 5 > gpurun_out/r02f_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_all.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --replicas 1 --repeats 5 > gpurun_out/r02f_ncu_launches.log 2>&1
python tools/show_bench.py gpurun_out/r02f_bench20.json
wc -l gpurun_out/r02f_launches_all.csv
