timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab6_tests.log 2>&1; tail -5 gpurun_out/ab6_tests.log
( time timeout 900 python bench.py > gpurun_out/ab6_bench.json 2>gpurun_out/ab6_bench.err ) 2>&1 | grep real; tail -3 gpurun_out/ab6_bench.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/ab6_ref.json 2>gpurun_out/ab6_ref.err ) 2>&1 | grep real
python - <<'PY'
import json
for f in ['gpurun_out/ab6_bench.json','gpurun_out/ab6_ref.json']:
  try:
    d=json.loads(open(f).read().strip().split('\n')[-1])
    print(f, '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])
    if 'roofline' in d:
        r=d['roofline']; print('  roofline', r['frac'], r['frac_actual'], r['issue_frac'], r['survivors_per_move'], r['evaluations_per_move'])
    if 'box' in d:
        b=d['box']; r=b['roofline']; print('  box %.4g'%b['value'], 'e2e %.4g'%b['e2e']['value'], b['ms_per_sweep'], r['frac'], r['frac_actual'], r['issue_frac'], r['survivors_per_move'], r['evaluations_per_move'])
    print('  cpu', d.get('cpu_baseline'))
  except Exception as e: print(f, 'FAILED', e)
PY
