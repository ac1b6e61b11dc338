mkdir -p gpurun_out
for r in CL FV LV LVW QS; do
  ( timeout 60 python tools/eqf_probe.py q_pow2_s29 $r > gpurun_out/r2_eqfp_$r.log 2>&1; echo "rc=$?" >> gpurun_out/r2_eqfp_$r.log ) 
  tail -4 gpurun_out/r2_eqfp_$r.log
done
for r in CL FV LV LVW QS; do
  if grep -q "rc=124" gpurun_out/r2_eqfp_$r.log; then
    GNE_EQF_DEBUG=0 timeout 60 python tools/eqf_probe.py q_pow2_s29 $r > /dev/null 2> gpurun_out/r2_eqfd_$r.err; tail -12 gpurun_out/r2_eqfd_$r.err; wc -l gpurun_out/r2_eqfd_$r.err
    tail -200 gpurun_out/r2_eqfd_$r.err > gpurun_out/r2_eqfd_$r.tail; rm gpurun_out/r2_eqfd_$r.err
  fi
done
