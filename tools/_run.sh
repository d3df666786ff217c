timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench20.log 2>&1; tail -1 gpurun_out/bench20.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['host_ms_per_step'], d['cpu_baseline'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench21.log 2>&1; tail -1 gpurun_out/bench21.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "ps_step/" -c 700 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --setup-steps 0 --no-cpu-baseline > gpurun_out/ncu_launch4.log 2>&1
