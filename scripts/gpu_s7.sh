#!/bin/bash
set -x
O=gpurun_out/s7
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --lanes 1 > $O/l1_heads.json 2> $O/l1_heads.err
BSW_TC_HEADS=0 $B --lanes 1 > $O/l1_noheads.json 2> $O/l1_noheads.err
$B --lanes 4 > $O/l4_heads.json 2> $O/l4_heads.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc$|k_conv_tc_head|k_conv_tc_h' -s 30 -c 8 -o $O/r2_convs python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > $O/ncu.log 2>&1
python scripts/overlap_probe.py 1024 60 > $O/ov1024_long.json 2> $O/ov1024_long.err
ls -la $O
