#!/bin/bash
# multi-GPU evidence run (gpurun --gpus N): the bench line with the NCCL gather inside the timed region, config 5 (crop)
N=${1:-2}
set -x
O=gpurun_out/multi$N
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_cifar8.json 2> $O/bench_cifar8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --config crop --steps 1 --warmup 1 > $O/bench_crop.json 2> $O/bench_crop.err
tail -3 $O/*.err
ls -la $O
