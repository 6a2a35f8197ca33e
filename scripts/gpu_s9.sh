#!/bin/bash
set -x
O=gpurun_out/s9
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $O/pytest.txt
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B --lanes 1 > $O/l1.json 2> $O/l1.err
$B --lanes 4 > $O/l4.json 2> $O/l4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows6|k_pop6|k_push_pairs' -s 8 -c 8 -o $O/r2_rows6q python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > $O/ncu_rows.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc$' -s 2 -c 4 -o $O/r2_conv5 python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > $O/ncu_conv5.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/bench_cifar8.json 2> $O/bench_cifar8.err
ls -la $O
