for w in 1 4 8; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --tc-waves $w > gpurun_out/bench14_$w.log 2>&1; tail -1 gpurun_out/bench14_$w.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('waves $w', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_per_step'], d['roofline']['all']['gemm_q_fwd_l0']['ms_per_step'], d['roofline']['all']['gemm_q_wgrad_l0']['ms_per_step'])"
done
