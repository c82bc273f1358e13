nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/ab5_tests.log 2>&1; tail -5 gpurun_out/ab5_tests.log
timeout 300 python bench.py --workload box --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/ab5_box1.json 2>gpurun_out/ab5_box1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload box --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/ab5_box2.json 2>gpurun_out/ab5_box2.err; tail -5 gpurun_out/ab5_box2.err
python - <<'PY'
import json
for f in ['gpurun_out/ab5_box1.json','gpurun_out/ab5_box2.json']:
  try:
    d=json.loads(open(f).read().strip().split('\n')[-1]); print(f, '%.4g'%d['value'], 'e2e', d['e2e'] and '%.4g'%d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['checks'])
  except Exception as e: print(f, 'FAILED', e)
PY
