# round-2 measurement record: run under gpurun from the repo root (one B200)
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2_final_gpu_tests.log 2>&1; tail -1 gpurun_out/r2_final_gpu_tests.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 400 gpurun_out/r2_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>&1
JPGB200_FUSED=1 python bench.py --frames 64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_final_bench_fused_k1e.json 2>/dev/null
# frame by frame (one 8K frame per launch, the state before the launch groups) beside the default
JPGB200_GROUP_LARGE=1 python bench.py --frames 64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-worst-case > gpurun_out/r2_final_bench_frame_by_frame.json 2>/dev/null
python profiles/run_other_configs.py > gpurun_out/r2_final_other_configs.log 2>&1
python profiles/run_noise_8k.py > gpurun_out/r2_final_noise_8k.log 2>&1
# 24 frames per step = three launch groups of eight 8K frames
CMD="python bench.py --frames 24 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-worst-case --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k "regex:k1_transform|k_encode_chunks|k_count_ff|k_scatter" --launch-skip 12 -c 8 -f -o gpurun_out/prof_r2_final $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
# one 8K frame per launch, for the per-frame traffic figures
CMD1="env JPGB200_GROUP_LARGE=1 python bench.py --frames 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-worst-case --no-extras"
ncu --set full --clock-control none --import-source on -k "regex:k1_transform|k_encode_chunks|k_count_ff|k_scatter" --launch-skip 8 -c 4 -f -o gpurun_out/prof_r2_final_single $CMD1 > gpurun_out/ncu_f1.log 2>&1
tail -2 gpurun_out/ncu_f1.log | cut -c1-200
