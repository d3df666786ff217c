timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test11.log
tail -4 gpurun_out/test11.log; grep -E "^E  " gpurun_out/test11.log | head
