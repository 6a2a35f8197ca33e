#!/bin/bash
set -x
O=gpurun_out/s5
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest.txt
python scripts/overlap_probe.py 256 12 > $O/ov256_base.json 2> $O/ov256_base.err
python scripts/overlap_probe.py 1024 4 > $O/ov1024_base.json 2> $O/ov1024_base.err
BSW_R6_PERSIST=148 python scripts/overlap_probe.py 1024 4 > $O/ov1024_p148.json 2> $O/ov1024_p148.err
BSW_TC_PERSIST=0 python scripts/overlap_probe.py 1024 4 > $O/ov1024_tile.json 2> $O/ov1024_tile.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --lanes 1 > $O/l1.json 2> $O/l1.err
$B --lanes 4 > $O/l4.json 2> $O/l4.err
for P in 2 1 0; do
  for L in 1 2 4; do
    BSW_CODEC_PRIO=$P $B --lanes $L --dual-stream 1 > $O/l${L}_dag_prio$P.json 2> $O/l${L}_dag_prio$P.err
  done
done
BSW_R6_PERSIST=148 BSW_CODEC_PRIO=0 $B --lanes 2 --dual-stream 1 > $O/l2_dag_prio0_p148.json 2> $O/l2_dag_prio0_p148.err
BSW_TC_PERSIST=0 BSW_CODEC_PRIO=0 $B --lanes 2 --dual-stream 1 > $O/l2_dag_prio0_tile.json 2> $O/l2_dag_prio0_tile.err
BSW_TC_PERSIST=0 BSW_CODEC_PRIO=0 $B --lanes 4 --dual-stream 1 > $O/l4_dag_prio0_tile.json 2> $O/l4_dag_prio0_tile.err
BSW_TC_PERSIST=0 $B --lanes 4 > $O/l4_tile.json 2> $O/l4_tile.err
ls -la $O
