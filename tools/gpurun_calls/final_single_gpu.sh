# round-2 final single-GPU evidence: the full GPU suite, the bench lines of every workload, the reference arm, the ncu passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_gputest_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputest_final.log
tail -3 gpurun_out/r2_gputest_final.log
timeout 200 python bench.py > gpurun_out/r2f_bench_c3_lv.json 2> gpurun_out/r2f_bench_c3_lv.err; tail -c 700 gpurun_out/r2f_bench_c3_lv.json
timeout 200 python bench.py --impl reference > gpurun_out/r2f_bench_c3_ref.json 2> gpurun_out/r2f_bench_c3_ref.err; tail -c 400 gpurun_out/r2f_bench_c3_ref.json
timeout 300 python bench.py --workload c5-lv --steps 50 --warmup 5 > gpurun_out/r2f_bench_c5_lv.json 2> gpurun_out/r2f_bench_c5_lv.err
timeout 120 python bench.py --workload c2-lv --steps 200 --warmup 20 > gpurun_out/r2f_bench_c2_lv.json 2> gpurun_out/r2f_bench_c2_lv.err
timeout 120 python bench.py --workload c2-qs --steps 200 --warmup 20 > gpurun_out/r2f_bench_c2_qs.json 2> gpurun_out/r2f_bench_c2_qs.err
timeout 120 python bench.py --workload c4-batch --batch 32 --steps 100 --warmup 10 > gpurun_out/r2f_bench_c4_batch32.json 2> gpurun_out/r2f_bench_c4_batch32.err
timeout 200 python tools/eqf_solve.py 3000 60000 35 lossy 20 3 LV --budget=20000 > gpurun_out/r2f_eqf_solve.txt 2>&1
timeout 200 python tools/eqf_solve.py 3000 60000 38 lossy 10 2 QS --budget=20000 >> gpurun_out/r2f_eqf_solve.txt 2>&1
cat gpurun_out/r2f_eqf_solve.txt
# ncu: launch list of the default command, then one full capture of the pricing kernel (numbers printed under ncu are not bench values)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_c3_lv.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_price_lv_stream -s 10 -c 2 -o gpurun_out/r2f_price_lv_stream -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_ncu_full.log 2>&1
ls -la gpurun_out/r2f_price_lv_stream.ncu-rep
