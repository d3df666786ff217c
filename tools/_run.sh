timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x > gpurun_out/test15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test15.log
tail -3 gpurun_out/test15.log; grep -E "^E  " gpurun_out/test15.log | head -5
