# ncu --set full capture of the tensor-core actor kernels (one launch each): run on the GPU box, read the .ncu-rep with
#   ncu -i gpurun_out/pol_fwd_tc.ncu-rep --page raw --csv      (profiles/r02_ncu_policy_tc_fwd_bwd.txt is the condensed form)
mkdir -p gpurun_out
for k in fwd_tc bwd_tc; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/pol_$k -f \
        python tools/policy_timing.py > gpurun_out/ncu_pol_$k.log 2>&1
    tail -2 gpurun_out/ncu_pol_$k.log
done
