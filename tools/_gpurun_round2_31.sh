mkdir -p gpurun_out
timeout 600 python tools/lbfgs_timing.py > gpurun_out/r2_lbfgs_timing.log 2>&1; tail -8 gpurun_out/r2_lbfgs_timing.log
