timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k sample_batch > gpurun_out/test6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test6.log
tail -4 gpurun_out/test6.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.log 2>&1; tail -1 gpurun_out/bench9.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('native', d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_per_step'])"
PS_NATIVE_SAMPLER=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.log 2>&1; tail -1 gpurun_out/bench10.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('torch', d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_per_step'])"
