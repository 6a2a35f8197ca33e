#!/bin/bash
# round-2 final single-GPU evidence run: ncu --set full of the conv kernels (raw page as CSV), the per-category facts
# bench.py's roofline reads (profiles/ncu_facts_r2.json, regenerated from that CSV and the tracked coder CSV), tests, smoke,
# the bench line, launch list, other configurations.  Outputs under gpurun_out/final/ (copied into profiles/ by hand).
# The coder kernels are unchanged since profiles/r2_final_rows_raw.csv was captured; pass "rows" to capture them again,
# "rowsonly" to capture only them (and skip the crop line), "all" to add the MNIST / HWC-quirk / reference-arm lines.
set -x
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi.txt
parts="convs:k_conv_tc:2:14"
[ "$1" = rows -o "$1" = all ] && parts="rows:k_rows6|k_pop6|k_push_pairs|k_rows<|k_pop_coarse:10:10 $parts"
[ "$1" = rowsonly ] && parts="rows:k_rows6|k_pop6|k_push_pairs|k_rows<|k_pop_coarse:10:10"
# ncu --set full: reports stay on the box when they are large (gpurun brings back at most 64 MiB); their raw page travels as CSV
for part in $parts; do
  IFS=: read name rx skip cnt <<< "$part"
  timeout 900 ncu --set full --clock-control none -k regex:"$rx" -s $skip -c $cnt -o /tmp/r2_final_$name python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > $O/ncu_$name.log 2>&1
  ncu -i /tmp/r2_final_$name.ncu-rep --page raw --csv > $O/r2_final_${name}_raw.csv 2>/dev/null
  [ $(stat -c %s /tmp/r2_final_$name.ncu-rep) -lt 20000000 ] && cp /tmp/r2_final_$name.ncu-rep $O/
  cp $O/r2_final_${name}_raw.csv profiles/
done
python scripts/ncu_summary.py facts 1024 profiles/ncu_facts_r2.json profiles/r2_final_rows_raw.csv profiles/r2_final_convs_raw.csv > $O/facts.log 2>&1
cp profiles/ncu_facts_r2.json $O/
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_cifar8.json 2> $O/bench_cifar8.err
python bench.py --steps 3 --warmup 3 --lanes 1 --no-cpu-baseline > $O/bench_cifar8_lanes1.json 2> $O/bench_cifar8_lanes1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 1000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline > $O/launches.log 2>&1
python bench.py --config imagenet4 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_imagenet4.json 2> $O/bench_imagenet4.err
[ "$1" = rowsonly ] || python bench.py --config crop --steps 1 --warmup 1 > $O/bench_crop_1gpu.json 2> $O/bench_crop_1gpu.err
if [ "$1" = all ]; then
  python bench.py --config mnist2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_mnist2.json 2> $O/bench_mnist2.err
  python bench.py --config crop --steps 1 --warmup 1 --crop-images 16 --hwc-quirk > $O/bench_crop_hwc_quirk_16.json 2> $O/bench_crop_hwc_quirk_16.err
  timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 2 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
fi
ls -la $O
