mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $T --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --workload c4-batch --batch 32 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2g_bench_c4_batch256_n8.json 2> gpurun_out/r2g_bench_c4_batch256_n8.err
tail -c 1500 gpurun_out/r2g_bench_c4_batch256_n8.json
