#!/bin/bash
# The round's single-GPU evidence in one gpurun call: full GPU test suite, the driver-style bench, ncu launch list and
# one --set full capture of the fused (8 k-blocks) bulk launch.  Every step runs under its own timeout.
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=12 --timeout 600 > $OUT/r02_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -18 $OUT/r02_pytest_full.log
timeout 600 python bench.py > $OUT/r02_bench_1gpu.json 2> $OUT/r02_bench_1gpu.err; echo "bench rc=$?"; tail -c 400 $OUT/r02_bench_1gpu.err
timeout 300 python bench.py --impl reference > $OUT/r02_bench_ref.json 2> $OUT/r02_bench_ref.err; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $OUT/r02_launches_n32768.csv python tools/one_solve.py 32768 > $OUT/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"; tail -2 $OUT/r02_ncu_list.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fw_bulk_kernel -s 271 -c 1 -o $OUT/r02_bulk_g8 -f python tools/one_solve.py 32768 > $OUT/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $OUT/r02_ncu_full.log
ncu -i $OUT/r02_bulk_g8.ncu-rep --page raw --csv > $OUT/r02_bulk_g8_raw.csv 2>/dev/null; wc -c $OUT/r02_bulk_g8_raw.csv
