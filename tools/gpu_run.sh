TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table 2>&1 | tail -1 | cut -c1-200
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table --no-broadcast-buffers 2>&1 | tail -1 | cut -c1-200
timeout 200 python bench.py --steps 10 --warmup 3 --no-kernel-table --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
