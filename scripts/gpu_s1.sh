#!/bin/bash
# round-2 GPU session 1: parity of the new affine-row kernels, A/B bench, ncu of k_rows6 / k_pop6
set -x
mkdir -p gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1/smi.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/s1/pytest.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1/bench_l4.json 2> gpurun_out/s1/bench_l4.err
python bench.py --steps 3 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/s1/bench_l1.json 2> gpurun_out/s1/bench_l1.err
BSW_ROWS_MODE=0 python bench.py --steps 3 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/s1/bench_l1_generic.json 2> gpurun_out/s1/bench_l1_generic.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows6|k_pop6' -s 6 -c 4 -o gpurun_out/s1/r2_rows6 python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > gpurun_out/s1/ncu.log 2>&1
ls -la gpurun_out/s1
