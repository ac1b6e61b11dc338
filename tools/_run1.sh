mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputest1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputest1.log
export GNE_BATCH_STATS=1
for cfg in "32 8 1" "32 8 2" "32 8 4" "32 8 8" "32 4 8" "32 0 8" "32 0 16"; do set -- $cfg
timeout 300 python bench.py --workload c4-batch --batch $1 --batch-group $2 --batch-threads $3 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_c4_b$1_g$2_t$3.json 2> gpurun_out/r2_c4_b$1_g$2_t$3.err
done
tail -3 gpurun_out/r2_gputest1.log
