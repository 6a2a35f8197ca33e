#!/bin/bash
# round-2 GPU session 3: parity of the persistent conv kernel + multi-row k_rows6, A/B benches, overlap probe with the persistent convs
set -x
O=gpurun_out/s3
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --lanes 1 > $O/l1_new.json 2> $O/l1_new.err
BSW_TC_PERSIST=0 $B --lanes 1 > $O/l1_tile.json 2> $O/l1_tile.err
BSW_R6_LPR=32 $B --lanes 1 > $O/l1_lpr32.json 2> $O/l1_lpr32.err
BSW_R6_LPR=8 $B --lanes 1 > $O/l1_lpr8.json 2> $O/l1_lpr8.err
BSW_R6_LPR=2 $B --lanes 1 > $O/l1_lpr2.json 2> $O/l1_lpr2.err
$B --lanes 4 > $O/l4_new.json 2> $O/l4_new.err
$B --lanes 2 > $O/l2_new.json 2> $O/l2_new.err
$B --lanes 4 --dual-stream 1 > $O/l4_ds1.json 2> $O/l4_ds1.err
$B --lanes 2 --dual-stream 1 > $O/l2_ds1.json 2> $O/l2_ds1.err
$B --lanes 4 --dual-stream 2 > $O/l4_ds2.json 2> $O/l4_ds2.err
python scripts/overlap_probe.py 256 12 > $O/overlap_256.json 2> $O/overlap_256.err
python scripts/overlap_probe.py 1024 4 > $O/overlap_1024.json 2> $O/overlap_1024.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows6|k_conv_tc_p' -s 6 -c 8 -o $O/r2_s3 python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > $O/ncu.log 2>&1
ls -la $O
