#!/bin/bash
# Round-end evidence on one B200: parity tests, bench (both arms), launch list, one ncu --set full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 400 gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_march|k_fill_columns' -s 4 -c 2 -f -o gpurun_out/r01_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
tail -2 gpurun_out/final_ncu.log
python __graft_entry__.py smoke
