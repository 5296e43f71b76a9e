mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2_pytest_kernels_12.log 2>&1; tail -5 gpurun_out/r2_pytest_kernels_12.log
for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-split --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | tee gpurun_out/r2_split_ab_$sz.log; done
