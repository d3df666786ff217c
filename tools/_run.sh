timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test13.log
tail -3 gpurun_out/test13.log; grep -E "^(FAILED|ERROR)" gpurun_out/test13.log | head
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.log 2>&1; tail -1 gpurun_out/bench17.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_per_step'], d['walk']['ms']); print({k:v['ms_per_step'] for k,v in d['roofline']['all'].items()})"
