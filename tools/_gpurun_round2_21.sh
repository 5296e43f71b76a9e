mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_igemm2_tf32_kernel -c 1 --profile-from-start off -o gpurun_out/r2_ncu_conv1_2_fwd python tools/profile_step.py --size 1080p --steps 1 > gpurun_out/ncu_c12.log 2>&1; tail -3 gpurun_out/ncu_c12.log
ls -la gpurun_out/r2_ncu_conv1_2_fwd.ncu-rep
