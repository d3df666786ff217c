for w in cfg1 cfg2; do
timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; tail -1 gpurun_out/bench_$w.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w ours', d['value'], d['ms_per_step'], d['e2e']['value'], d['host_ms_per_step'])"
timeout 300 python bench.py --workload $w --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$w.log 2>&1; tail -1 gpurun_out/bench_ref_$w.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w reference-port', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
done
