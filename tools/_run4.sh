mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_eqf.py -q -m gpu > gpurun_out/r2_eqf2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_eqf2.log
timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "equal_flow" >> gpurun_out/r2_eqf2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_eqf2.log
tail -15 gpurun_out/r2_eqf2.log
timeout 300 python tools/eqf_solve.py 3000 60000 35 lossy 20 3 LV --budget=20000 > gpurun_out/r2_eqf_solve.txt 2>&1
timeout 300 python tools/eqf_solve.py 3000 60000 38 lossy 10 2 QS --budget=20000 >> gpurun_out/r2_eqf_solve.txt 2>&1
cat gpurun_out/r2_eqf_solve.txt
