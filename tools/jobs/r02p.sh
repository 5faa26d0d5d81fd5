set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_rollout_gpu.py tests/test_rollout_graph_gpu.py -q -x > gpurun_out/r02p_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r02p_pytest.txt
tail -4 gpurun_out/r02p_pytest.txt
for c in 10 25 50; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --e2e-chunk $c > gpurun_out/r02p_b20_c$c.json 2> gpurun_out/r02p_b20_c$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_b20_c*.json')):
    d=json.loads([l for l in open(f) if l.startswith('{')][0]); print(f, 'value %.4g e2e %.4g ms/step %.5f closed %.4g' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['closed_loop']['value']))
PY
