python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2>/dev/null
python bench.py --workload chains --precision mixed --no-cpu-baseline > gpurun_out/r02_bench_mixed.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --equil 10 > gpurun_out/r02_launches.log 2>&1
PMC_SPEC_QUEUE=0 ncu --set full --clock-control none --import-source on -k regex:k_chain_sweep_spec -s 3 -c 1 -o gpurun_out/r02_final_spec python bench.py --workload chains --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sweeps 1 --equil 20 > gpurun_out/r02_final_spec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_box_sweep_all -s 6 -c 1 -o gpurun_out/r02_final_box python bench.py --workload box --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sweeps 1 --equil 5 > gpurun_out/r02_final_box.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
