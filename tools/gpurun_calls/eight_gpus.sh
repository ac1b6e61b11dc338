# one 8-GPU box: sharded parity tests on real NVLink peers, configs[2] / configs[4] sharded over 8, configs[3] as 256 instances
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
timeout 240 python -m pytest tests/test_multi_rank.py -q -m gpu -k "fused_exchange or shorter_than or c2_prefix or first_violator" > gpurun_out/r2_multirank_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multirank_8gpu.log
timeout 150 $T --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_c3_lv_n8.json 2> gpurun_out/r2_bench_c3_lv_n8.err
timeout 240 $T --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --workload c5-lv --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c5_lv_n8.json 2> gpurun_out/r2_bench_c5_lv_n8.err
timeout 200 $T --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --workload c4-batch --batch 32 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_c4_batch256_n8.json 2> gpurun_out/r2_bench_c4_batch256_n8.err
timeout 120 $T --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_c3_lv_n4.json 2> gpurun_out/r2_bench_c3_lv_n4.err
timeout 120 $T --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_c3_lv_n2.json 2> gpurun_out/r2_bench_c3_lv_n2.err
tail -3 gpurun_out/r2_multirank_8gpu.log; for f in gpurun_out/r2_bench_c3_lv_n8 gpurun_out/r2_bench_c5_lv_n8 gpurun_out/r2_bench_c4_batch256_n8; do tail -c 600 $f.json; tail -2 $f.err; done
