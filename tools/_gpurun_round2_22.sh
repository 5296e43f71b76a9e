mkdir -p gpurun_out
run() { for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-resident --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | grep "round 1 rule" | sed "s/^/$1 $sz: /"; done; }
run "no hint     "
for ns in 2000 300; do
STV_NVCC_EXTRA="-DSTV_TRYWAIT_HINT_NS=$ns" python build_native.py --force > /dev/null 2>&1; echo "build rc=$?"
run "hint ${ns} ns"
done
python build_native.py --force > /dev/null 2>&1
run "no hint again"
