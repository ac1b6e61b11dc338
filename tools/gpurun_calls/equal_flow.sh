mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_eqf.py tests/test_gpu_parity.py -q -m gpu -k "eqf or equal_flow" > gpurun_out/r2_eqf6.log 2>&1; echo rc=$? >> gpurun_out/r2_eqf6.log; tail -25 gpurun_out/r2_eqf6.log
timeout 200 python tools/eqf_solve.py 3000 60000 35 lossy 20 3 LV --budget=20000 > gpurun_out/r2h_eqf_solve.txt 2>&1
timeout 200 python tools/eqf_solve.py 3000 60000 38 lossy 10 2 QS --budget=20000 >> gpurun_out/r2h_eqf_solve.txt 2>&1
cat gpurun_out/r2h_eqf_solve.txt
