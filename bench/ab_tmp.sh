nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_box.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload box --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/ab21_box1.json 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload box --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/ab21_box2.json 2>/dev/null
python - <<'PY'
import json
for f in ['gpurun_out/ab21_box1.json','gpurun_out/ab21_box2.json']:
    d=json.loads(open(f).read().strip().split('\n')[-1]); print(f, '%.4g'%d['value'], d['ms_per_step'])
PY
