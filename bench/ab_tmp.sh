timeout 900 python -m pytest tests/test_gpu_box.py tests/test_gpu_multirank.py tests/test_gpu_observables.py tests/test_gpu_edge.py -x -q 2>&1 | tail -3
for v in prev default; do
  lib=libpmc_b200_$v.so; [ $v = default ] && lib=libpmc_b200.so
  for n in 1048576 131072; do
  PMC_B200_LIB=$lib timeout 300 python bench.py --workload box --box-particles $n --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/ab19_${v}_$n.json 2>/dev/null
  python - gpurun_out/ab19_${v}_$n.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], 'box %.4g ms/step %.4f'%(d['value'], d['ms_per_step']), d['checks']['acceptance'])
PY
  done
done
