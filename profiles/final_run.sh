set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --frames 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-worst-case > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_transform|k_encode|k_count|k_scatter" --launch-skip 8 -c 4 -f -o gpurun_out/prof_r1_final python bench.py --frames 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-worst-case > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
