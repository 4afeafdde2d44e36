mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bwd_tc -c 1 -o gpurun_out/pol_bwd_tc -f python tools/policy_timing.py > gpurun_out/ncu_pol.log 2>&1
tail -3 gpurun_out/ncu_pol.log
