# end-of-round measurement run (one B200): tests, bench lines, ncu launch list and --set full captures -> gpurun_out/
# usage: bash tests/gpu_final_run.sh TAG        (profiles/summarise.py turns the captures into profiles/TAG_*)
TAG=${1:-final}
set -x
timeout 1100 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench_config2_n1.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_config2_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --config 5 --steps 3 --warmup 2 --no-cpu > gpurun_out/${TAG}_bench_config5_one_rank.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > gpurun_out/${TAG}_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_flood2|k_tile_front|k_rag_accumulate|k_agglomerate_par|k_mask_bits_u8|k_finalize|k_fragstats|k_crop_union|k_relabel_dense" -c 9 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-extra > gpurun_out/${TAG}_ncu_f.log 2>&1
ls -la gpurun_out/${TAG}_prof.ncu-rep
BS_MWS_VERBOSE=1 timeout 400 python tests/gpu_mws_scale.py 256 512 512 > gpurun_out/${TAG}_mws_scale.jsonl 2>&1; tail -2 gpurun_out/${TAG}_mws_scale.jsonl | cut -c1-300
