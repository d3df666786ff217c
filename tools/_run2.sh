for ex in "" "--exchange"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --mode infer --workload cfg4q --steps 2 --warmup 1 $ex > gpurun_out/infer_cfg4q_2gpu$ex.log 2>&1; echo "rc=$?" >> gpurun_out/infer_cfg4q_2gpu$ex.log
grep '^{"metric' gpurun_out/infer_cfg4q_2gpu$ex.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$ex', d['value'], d['ms_per_pass'], d['checksum_rank0'], d['config']['rank0_closure'])"
done
