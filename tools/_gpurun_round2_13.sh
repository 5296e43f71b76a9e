mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_first_fwd_kernel -c 1 --profile-from-start off -o gpurun_out/r2_ncu_conv_first_v3 python tools/profile_step.py --size 1080p --steps 1 > gpurun_out/ncu_first.log 2>&1; tail -3 gpurun_out/ncu_first.log
ls -la gpurun_out/r2_ncu_conv_first_v3.ncu-rep
