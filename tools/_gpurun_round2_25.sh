mkdir -p gpurun_out
for c in conv_compact conv_variants conv2; do timeout 600 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_${c}_25.log 2>&1; echo "$c rc=$? pass=$(grep -c ^PASS gpurun_out/r2_selftest_${c}_25.log) fail=$(grep -c ^FAIL gpurun_out/r2_selftest_${c}_25.log)"; grep ^FAIL gpurun_out/r2_selftest_${c}_25.log | head -8; done
for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-pool --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | tee gpurun_out/r2_pool_ab_$sz.log | tail -4; done
