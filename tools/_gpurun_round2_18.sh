mkdir -p gpurun_out
for sz in 512 1080p; do
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex.sum,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${sz}_step_v3.csv python tools/profile_step.py --size $sz --steps 1 > gpurun_out/ncu_${sz}.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_${sz}_step_v3.csv | head -22
done
