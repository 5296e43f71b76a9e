mkdir -p gpurun_out
run() { for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-resident --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | grep "round 1 rule" | sed "s/^/$1 $sz: /"; done; }
for c in conv_compact conv_variants; do timeout 600 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_${c}_23.log 2>&1; echo "$c rc=$? pass=$(grep -c ^PASS gpurun_out/r2_selftest_${c}_23.log) fail=$(grep -c ^FAIL gpurun_out/r2_selftest_${c}_23.log)"; done
run "new pool epilogue"
STV_NVCC_EXTRA="-DSTV_AB_OLD_POOL" python build_native.py --force > /dev/null 2>&1; echo "build rc=$?"
run "old pool epilogue"
