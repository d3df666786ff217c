timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_d.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu_d.log
grep '^{"metric' gpurun_out/bench_2gpu_d.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2gpu', d['value'], d['ms_per_step'], d['e2e'], d['host_ms_per_step'])"
tail -3 gpurun_out/bench_2gpu_d.log | cut -c1-300
