#!/bin/bash
# one GPU-box pass: tests, bench, launch list, full ncu capture of the dominant kernel (outputs under gpurun_out/)
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err; echo "ref exit $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/launches_${TAG}_cfg1.csv python bench.py --workload cfg1 --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launch_cfg1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mode_sum_kernel -s 4 -c 1 -o gpurun_out/${TAG}_modesum -f \
  python bench.py --batch 16 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
