mkdir -p gpurun_out
timeout 200 python bench.py --workload c4-batch --batch 32 --steps 200 --warmup 10 > gpurun_out/r2g_bench_c4_batch32.json 2> gpurun_out/r2g_bench_c4_batch32.err
tail -c 1500 gpurun_out/r2g_bench_c4_batch32.json
