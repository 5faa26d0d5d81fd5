set -x
cd $GRAFT_REPO_ROOT
bash tools/jobs/ab.sh r02m v1 t1
timeout 300 python tools/lone_launch.py > gpurun_out/r02m_lone_pdl.txt 2>&1
GLG_GRAPH_PDL=0 timeout 300 python tools/lone_launch.py > gpurun_out/r02m_lone_nopdl.txt 2>&1
cat gpurun_out/r02m_lone_pdl.txt gpurun_out/r02m_lone_nopdl.txt
