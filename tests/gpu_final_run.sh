# end-of-round measurement run (one B200): tests, bench lines, ncu launch list and --set full captures -> gpurun_out/
set -x
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_flood2|k_rag_accumulate|k_agglomerate_par" -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/prof_final.ncu-rep
