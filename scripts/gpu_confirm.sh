#!/bin/bash
# confirmation run of the final tree: tests, smoke, the bench line (with cpu_baseline and ncu facts), the reference arm as the driver launches it
set -x
O=gpurun_out/confirm
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/bench_cifar8.json 2> $O/bench_cifar8.err
ls -la $O
