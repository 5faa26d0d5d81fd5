set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02q_bench20_n$N.json 2> gpurun_out/r02q_bench20_n$N.err; echo "rc=$?" >> gpurun_out/r02q_bench20_n$N.err
tail -5 gpurun_out/r02q_bench20_n$N.err
python tools/show_bench.py gpurun_out/r02q_bench20_n$N.json
