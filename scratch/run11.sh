cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r11_pytest.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r11_pytest.log
timeout 500 python bench.py --workload cv4abc > gpurun_out/r11_cv.json 2> gpurun_out/r11_cv.err; echo "cv rc $?"; tail -3 gpurun_out/r11_cv.err; cat gpurun_out/r11_cv.json
