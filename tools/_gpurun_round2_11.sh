# 64-wide two-half tiles with the style accumulator; plan sweep of the first layers at 1080p / 512
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "split or bit_for_bit or conv_compact" > gpurun_out/r2_pytest_kernels_11.log 2>&1; tail -5 gpurun_out/r2_pytest_kernels_11.log
timeout 900 python tools/plan_sweep.py --size 1080p --layers 0,1,2,3 --reps 30 > gpurun_out/r2_plan_sweep_1080p_0123.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r2_plan_sweep_1080p_0123.log
timeout 600 python tools/plan_sweep.py --size 512 --layers 0,1,2,3 --reps 150 > gpurun_out/r2_plan_sweep_512_0123b.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r2_plan_sweep_512_0123b.log
