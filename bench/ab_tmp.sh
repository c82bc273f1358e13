timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_simulation.py tests/test_gpu_edge.py tests/test_gpu_observables.py -x -q 2>&1 | tail -15
python bench/other_configs.py > gpurun_out/ab13_other.json 2>gpurun_out/ab13_other.err; tail -3 gpurun_out/ab13_other.err
python - <<'PY'
import json
for line in open('gpurun_out/ab13_other.json'):
    d=json.loads(line); print('%.4g'%d['value'], d['config'][:90], d['acceptance_per_move'], d['energy_bookkeeping_rel_drift'])
PY
