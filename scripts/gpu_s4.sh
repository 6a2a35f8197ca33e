#!/bin/bash
set -x
O=gpurun_out/s4
mkdir -p $O
python scripts/debug_rows6.py > $O/debug_rows6.txt 2>&1
python -m pytest tests/test_codec_gpu.py -x -q -k "overlap_mode or persistent_kernel" 2>&1 | tail -15 > $O/pytest_new.txt
python scripts/overlap_probe.py 256 12 > $O/ov256_base.json 2> $O/ov256_base.err
BSW_R6_PERSIST=148 python scripts/overlap_probe.py 256 12 > $O/ov256_p148.json 2> $O/ov256_p148.err
BSW_R6_PERSIST=148 python scripts/overlap_probe.py 1024 4 > $O/ov1024_p148.json 2> $O/ov1024_p148.err
BSW_R6_PERSIST=296 python scripts/overlap_probe.py 1024 4 > $O/ov1024_p296.json 2> $O/ov1024_p296.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --lanes 1 --dual-stream 1 > $O/l1_dag.json 2> $O/l1_dag.err
$B --lanes 2 --dual-stream 1 > $O/l2_dag.json 2> $O/l2_dag.err
$B --lanes 4 --dual-stream 1 > $O/l4_dag.json 2> $O/l4_dag.err
BSW_R6_PERSIST=148 $B --lanes 1 --dual-stream 1 > $O/l1_dag_p148.json 2> $O/l1_dag_p148.err
BSW_R6_PERSIST=148 $B --lanes 2 --dual-stream 1 > $O/l2_dag_p148.json 2> $O/l2_dag_p148.err
BSW_R6_PERSIST=148 $B --lanes 4 --dual-stream 1 > $O/l4_dag_p148.json 2> $O/l4_dag_p148.err
BSW_R6_PERSIST=148 $B --lanes 1 > $O/l1_p148.json 2> $O/l1_p148.err
BSW_TC_PERSIST=0 $B --lanes 4 > $O/l4_tile.json 2> $O/l4_tile.err
BSW_TC_PERSIST=0 $B --lanes 2 --dual-stream 1 > $O/l2_dag_tile.json 2> $O/l2_dag_tile.err
ls -la $O
