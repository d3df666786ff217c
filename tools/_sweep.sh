for rep in 1 2 3; do for f in 1 2 0; do
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --steps-in-flight $f 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight',$f,'value_ms',l['ms_per_step'],'e2e',l['e2e']['ms_per_step'],l['e2e']['ms_per_step_median'],l['e2e']['ms_per_step_max'], l['host_ms_per_step'], l['cudaMallocs_in_timed_region'])"
done; done
