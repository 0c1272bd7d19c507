#!/bin/bash
# Standard GPU-box sequence: parity tests, bench, ncu launch list, ncu full capture of the
# dominant kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python scripts/kprof.py > gpurun_out/${TAG}_kprof.txt 2>&1; cat gpurun_out/${TAG}_kprof.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cat gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-bam --e2e-steps 1 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_project|k_pileup_gather|k_ctg_phase|k_scan_records" -s 8 -c 4 -f -o gpurun_out/${TAG}_prof4 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-bam --e2e-steps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -15
