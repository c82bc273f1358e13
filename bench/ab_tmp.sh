timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed.py -x -q 2>&1 | tail -3
B="python bench.py --workload chains --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
for rep in 1 2; do for v in prev default; do
  lib=libpmc_b200_$v.so; [ $v = default ] && lib=libpmc_b200.so
  PMC_B200_LIB=$lib $B > gpurun_out/ab18_${v}_$rep.json 2>/dev/null
  PMC_B200_LIB=$lib $B --precision mixed > gpurun_out/ab18_${v}_mixed_$rep.json 2>/dev/null
done; done
for f in gpurun_out/ab18_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1]); print(sys.argv[1], '%.4g'%d['value'])
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
