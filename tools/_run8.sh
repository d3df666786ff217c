timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_8gpu.log
grep '^{"metric' gpurun_out/bench_8gpu.log | cut -c1-260
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --mode infer --workload cfg4 --steps 2 --warmup 1 > gpurun_out/infer_cfg4_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/infer_cfg4_8gpu.log
grep '^{"metric' gpurun_out/infer_cfg4_8gpu.log | cut -c1-900
