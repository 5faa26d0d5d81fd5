set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02g_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.txt
tail -15 gpurun_out/r02g_pytest.txt
for c in 5 10 25 50; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --e2e-chunk $c > gpurun_out/r02g_bench20_c$c.json 2> gpurun_out/r02g_bench20_c$c.err
done
python tools/show_bench.py gpurun_out/r02g_bench20_c*.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_bench20_c*.json')):
    d=json.loads([l for l in open(f) if l.startswith('{')][0]); print(f, d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['closed_loop']['value'])
PY
