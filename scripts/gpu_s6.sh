#!/bin/bash
set -x
O=gpurun_out/s6
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --lanes 1 > $O/l1_tile.json 2> $O/l1_tile.err
BSW_TC_PERSIST=7 $B --lanes 1 > $O/l1_p7.json 2> $O/l1_p7.err
for S in 0 2000 4000 8000; do
  $B --lanes 4 --stagger-us $S > $O/l4_free_s$S.json 2> $O/l4_free_s$S.err
done
$B --lanes 4 --free-running 0 --stagger-us 0 > $O/l4_joined.json 2> $O/l4_joined.err
$B --lanes 2 --stagger-us 8000 > $O/l2_free_s8000.json 2> $O/l2_free_s8000.err
$B --lanes 3 --stagger-us 5000 > $O/l3_free_s5000.json 2> $O/l3_free_s5000.err
$B --lanes 8 --stagger-us 2000 > $O/l8_free_s2000.json 2> $O/l8_free_s2000.err
BSW_TC_PERSIST=1 $B --lanes 4 --stagger-us 4000 > $O/l4_free_s4000_p1.json 2> $O/l4_free_s4000_p1.err
$B --lanes 4 --stagger-us 4000 --dual-stream 1 > $O/l4_free_s4000_dag.json 2> $O/l4_free_s4000_dag.err
ls -la $O
