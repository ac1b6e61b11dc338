mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batch_driver or col_file_loader" > gpurun_out/r2_t10.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t10.log; tail -4 gpurun_out/r2_t10.log
export GNE_BATCH_STATS=1
for t in 1 2 4 8; do
timeout 200 python bench.py --workload c4-batch --batch 32 --batch-threads $t --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_c4a_t$t.json 2> gpurun_out/r2_c4a_t$t.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_c4a_t$t.json").read().strip().splitlines()[-1])
print("threads $t: e2e %.1f k pivots/s, %.1f us per round; device leg %.1f us per round, frac %.3f" % (d["e2e"]["pivots_per_s"]/1e3, 1e3*d["e2e"]["ms_per_pivot_per_instance"], 1e3*d["ms_per_step"], d["roofline"]["frac"]))
PY
done
