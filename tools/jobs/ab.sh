# usage: bash tools/jobs/ab.sh tag variant1 variant2 ...   (variants built by tools/ab.py into tools/ab/)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
for v in "$@"; do
  export GLG_LIB_PATH=$GRAFT_REPO_ROOT/tools/ab/libglg_$v.so
  timeout 600 python -m pytest tests/test_fused_rollout_gpu.py -q -x > gpurun_out/${TAG}_${v}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_${v}_pytest.txt
  tail -2 gpurun_out/${TAG}_${v}_pytest.txt
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_${v}_b20.json 2> gpurun_out/${TAG}_${v}_b20.err
  timeout 300 python bench.py --steps 2000 --warmup 100 --no-cpu --no-extra > gpurun_out/${TAG}_${v}_b2000.json 2> gpurun_out/${TAG}_${v}_b2000.err
done
unset GLG_LIB_PATH
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*_b*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print('%-40s value %.4g  us/step %.3f  frac %.4f  e2e %.3g  parity %s' % (f.split('/')[-1], d['value'], 1e3*d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity']['mismatches']))
    except Exception as e: print(f, 'ERR', e)
PY
