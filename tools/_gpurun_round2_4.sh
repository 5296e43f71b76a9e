mkdir -p gpurun_out
python tools/gpu_selftest.py --case halo > gpurun_out/r2_selftest_halo.log 2>&1; tail -14 gpurun_out/r2_selftest_halo.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "non_finite or release or lbfgs" > gpurun_out/r2_pytest_gpu_4.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu_4.log
timeout 600 python tools/gpu_selftest.py --case conv_variants > gpurun_out/r2_selftest_variants.log 2>&1; grep -c PASS gpurun_out/r2_selftest_variants.log; grep FAIL gpurun_out/r2_selftest_variants.log | head
for sz in 512 1080p; do
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${sz}_step.csv python tools/profile_step.py --size $sz --steps 1 > gpurun_out/ncu_${sz}.log 2>&1
done
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"conv_igemm2|gram_partial|conv_first" --csv --log-file gpurun_out/r2_ncu_conv_gram_metrics_1080p.csv python tools/profile_step.py --size 1080p --steps 1 > gpurun_out/ncu_m1080.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"conv_igemm2|gram_partial|conv_first" --csv --log-file gpurun_out/r2_ncu_conv_gram_metrics_512.csv python tools/profile_step.py --size 512 --steps 1 > gpurun_out/ncu_m512.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_1080p_step.csv | head -30
python tools/summarize_launches.py gpurun_out/r2_launches_512_step.csv | head -30
python tools/summarize_metrics.py gpurun_out/r2_ncu_conv_gram_metrics_1080p.csv
python tools/summarize_metrics.py gpurun_out/r2_ncu_conv_gram_metrics_512.csv
