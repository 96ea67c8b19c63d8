#!/bin/bash
# Round-2 multi-GPU pass (one 8-GPU box, gpurun --gpus 8): weak-scaling of the generate bench at N = 8, the config-4 training
# step at N = 2 / 4 / 8 (bucketed NCCL all-reduce over NVLink), and the second-device handle test.
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_config_parity_gpu.py -m gpu -q -k second_device 2>&1 | tail -2 > $O/r2_second_device_test.log
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --config train --steps 10 --warmup 3 > $O/r2_bench_train_n$n.json 2> $O/r2_bench_train_n$n.err
done
NCCL_DEBUG=INFO $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --config train --steps 3 --warmup 3 2>&1 | grep -E "NVLS|Channel|Ring|Tree|nranks" | head -12 > $O/r2_nccl_info_train_n8.log
$TR --nproc-per-node 8 --master-port 29530 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
tail -c 400 $O/r2_bench_train_n8.json; echo; tail -c 300 $O/r2_bench_n8.json; echo; cat $O/r2_second_device_test.log
