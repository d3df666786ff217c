timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu_b.log 2>&1; echo "rc=$?" >> gpurun_out/bench_8gpu_b.log
grep '^{"metric' gpurun_out/bench_8gpu_b.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('8gpu', d['value'], d['ms_per_step'], d['e2e'], d['host_ms_per_step'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_4gpu_b.log 2>&1; echo "rc=$?" >> gpurun_out/bench_4gpu_b.log
grep '^{"metric' gpurun_out/bench_4gpu_b.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('4gpu', d['value'], d['ms_per_step'], d['e2e'], d['host_ms_per_step'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --mode infer --workload cfg4 --steps 2 --warmup 1 --exchange > gpurun_out/infer_cfg4_8gpu_ex.log 2>&1; echo "rc=$?" >> gpurun_out/infer_cfg4_8gpu_ex.log
grep '^{"metric' gpurun_out/infer_cfg4_8gpu_ex.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 8gpu exchange', d['value'], d['ms_per_pass'], d['checksum_rank0'], d['hbm_gb_allocated'])"
