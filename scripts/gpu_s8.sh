#!/bin/bash
set -x
O=gpurun_out/s8
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --lanes 1 > $O/l1.json 2> $O/l1.err
$B --lanes 4 > $O/l4.json 2> $O/l4.err
$B --lane-size 296 > $O/ls296.json 2> $O/ls296.err
$B --lane-size 370 > $O/ls370.json 2> $O/ls370.err
$B --lane-size 222 > $O/ls222.json 2> $O/ls222.err
$B --lanes 3 > $O/l3.json 2> $O/l3.err
BSW_TC_PERSIST=1 $B --lanes 4 > $O/l4_p1.json 2> $O/l4_p1.err
python bench.py --config crop --steps 1 --warmup 1 > $O/crop_l4.json 2> $O/crop_l4.err
python bench.py --config crop --steps 1 --warmup 1 --lanes 8 > $O/crop_l8.json 2> $O/crop_l8.err
python bench.py --config imagenet4 --steps 2 --warmup 3 --no-cpu-baseline > $O/imagenet4.json 2> $O/imagenet4.err
ls -la $O
