TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for L in 4 2 1 8; do
  echo "bucket layers $L"
  MGV_TRAIN_BUCKET_LAYERS=$L $TR --nproc-per-node 8 --master-port $((29600 + L)) bench.py --gpus 8 --config train --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
