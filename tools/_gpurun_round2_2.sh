mkdir -p gpurun_out
python tools/grad_diagnose.py 1080 1920 --compact 1 > gpurun_out/r2_grad_diag_1080_c1.log 2>&1; cat gpurun_out/r2_grad_diag_1080_c1.log | tail -30
python tools/grad_diagnose.py 1080 1920 --compact 0 > gpurun_out/r2_grad_diag_1080_c0.log 2>&1; grep "rel L2 full\|worst" gpurun_out/r2_grad_diag_1080_c0.log
python tools/grad_diagnose.py 270 480 --compact 1 > gpurun_out/r2_grad_diag_270_c1.log 2>&1; grep "rel L2 full\|worst\|parity" gpurun_out/r2_grad_diag_270_c1.log
timeout 1500 python -m pytest tests -m gpu -q --deselect "tests/test_gpu_parity.py::test_adam_case_matches_reference" > gpurun_out/r2_pytest_gpu_2.log 2>&1; tail -15 gpurun_out/r2_pytest_gpu_2.log
