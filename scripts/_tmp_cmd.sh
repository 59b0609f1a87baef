python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches_bench_f64.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frame-check > gpurun_out/r02n_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/r02n_cfg3_f64 -f python scripts/gpu_one.py cfg3_cornell_1080p_4spp_d5 f64 3 > gpurun_out/r02n_ncu2.log 2>&1
NT_ONE_SHARDS=8 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/r02n_cfg3_f64_shard8 -f python scripts/gpu_one.py cfg3_cornell_1080p_4spp_d5 f64 3 > gpurun_out/r02n_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
