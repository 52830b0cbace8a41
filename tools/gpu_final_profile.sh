#!/bin/bash
# Round evidence on one B200: parity tests, bench (both arms), launch list, ncu --set full captures of the hot kernels.
# usage: tools/gpu_final_profile.sh <tag>
tag=${1:-r02_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; tail -c 300 gpurun_out/${tag}_bench_reference.json; echo
python tools/sweep_kernel_bench.py cfg3 tma; python tools/sweep_kernel_bench.py cfg3 reg
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-general-paths > gpurun_out/${tag}_launches.log 2>&1
# full-image march launches: 3 warm-up + timed ones; fill launches likewise: capture the 4th of each
ncu --set full --clock-control none --import-source on -k regex:'k_march_flat|k_fill_columns' -s 6 -c 2 -f -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-general-paths > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
ncu --set full --clock-control none --import-source on -k regex:'k_sweep_tma' -s 2 -c 1 -f -o gpurun_out/${tag}_sweep python tools/sweep_kernel_bench.py cfg3 tma > gpurun_out/${tag}_ncu_sweep.log 2>&1
tail -2 gpurun_out/${tag}_ncu_sweep.log
python __graft_entry__.py smoke
