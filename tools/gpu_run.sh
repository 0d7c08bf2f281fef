TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-330
$TR bench.py --gpus 2 --steps 10 --warmup 3 --ddp-bucket-mb 25 --no-kernel-table 2>&1 | tail -1 | cut -c1-200
$TR bench.py --gpus 2 --steps 10 --warmup 3 --sync-iqbn --no-kernel-table 2>&1 | tail -1 | cut -c1-200
