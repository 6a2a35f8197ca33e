#!/bin/bash
# round-2 GPU session 2: parity, new bench line (cpu baseline, trimmed e2e), crop config, overlap probe, ncu facts
set -x
O=gpurun_out/s2
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest.txt
python bench.py --steps 3 --warmup 3 > $O/bench_l4.json 2> $O/bench_l4.err
python bench.py --steps 3 --warmup 3 --lanes 2 --no-cpu-baseline > $O/bench_l2.json 2> $O/bench_l2.err
python bench.py --config crop --steps 1 --warmup 1 > $O/bench_crop.json 2> $O/bench_crop.err
python scripts/overlap_probe.py 256 12 > $O/overlap_256.json 2> $O/overlap_256.err
BSW_R6_WARPS=12 python scripts/overlap_probe.py 256 12 > $O/overlap_256_w12.json 2> $O/overlap_256_w12.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows6|k_pop6|k_push_pairs' -s 8 -c 5 -o $O/r2_rows6b python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > $O/ncu.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
ls -la $O
