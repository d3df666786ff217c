for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench23.log 2>&1; tail -1 gpurun_out/bench23.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('run $i', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_per_step'], d['clocks'])"
done
