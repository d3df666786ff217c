timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/test8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test8.log
tail -3 gpurun_out/test8.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.log 2>&1; tail -1 gpurun_out/bench12.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_per_step']); print({k:v['ms_per_step'] for k,v in d['roofline']['all'].items() if 'aggregate' in k})"
