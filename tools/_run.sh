timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -q -x > gpurun_out/test14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test14.log
tail -3 gpurun_out/test14.log; grep -E "^E  " gpurun_out/test14.log | head -5
timeout 300 python tools/knn_microbench.py > gpurun_out/knn_mb.log 2>&1; cat gpurun_out/knn_mb.log
