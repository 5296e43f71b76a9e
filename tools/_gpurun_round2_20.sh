mkdir -p gpurun_out
timeout 1500 python tools/gpu_parity_report.py --lbfgs > gpurun_out/r2_parity_report_v3.log 2>&1; echo "rc=$?"; grep -c "graph=" gpurun_out/r2_parity_report_v3.log; grep -A2 "lbfgs_noisy" gpurun_out/r2_parity_report_v3.log | head
