mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export GNE_BATCH_STATS=1
timeout 200 $T --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --workload c4-batch --batch 32 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_c4_batch256_n8_t1.json 2> gpurun_out/r2_bench_c4_batch256_n8_t1.err
timeout 200 $T --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --workload c4-batch --batch 32 --batch-threads 2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_c4_batch256_n8_t2.json 2> gpurun_out/r2_bench_c4_batch256_n8_t2.err
for f in gpurun_out/r2_bench_c4_batch256_n8_t1 gpurun_out/r2_bench_c4_batch256_n8_t2; do tail -c 900 $f.json; grep "batch thread" $f.err | tail -3; done
