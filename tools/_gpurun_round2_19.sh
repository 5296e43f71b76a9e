mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "stationary or split" > gpurun_out/r2_pytest_kernels_19.log 2>&1; tail -3 gpurun_out/r2_pytest_kernels_19.log
for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-resident --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | tee gpurun_out/r2_resident_ab_$sz.log; done
