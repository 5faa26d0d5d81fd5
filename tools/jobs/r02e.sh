set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02e_bench20_full.json 2> gpurun_out/r02e_bench20_full.err; echo "rc=$?" >> gpurun_out/r02e_bench20_full.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02e_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02e_ncu_launches.log 2>&1
timeout 300 python tools/pacman_perf.py > gpurun_out/r02e_pacman.txt 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pacman_observe -s 12 -c 1 -o gpurun_out/prof_r02e_pacman -f python tools/pacman_perf.py > gpurun_out/r02e_ncu_pacman.log 2>&1
python tools/show_bench.py gpurun_out/r02e_bench20_full.json
cat gpurun_out/r02e_pacman.txt
