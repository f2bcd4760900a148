# usage: ab_libs.sh tag lib1 lib2 ...   (libs in profiles/ab/libs; "default" = the product build)
tag=$1; shift
cp jpeg-encoder-and-decoder_b200/libjpgb200.so /tmp/default.so
for l in "$@"; do
  if [ "$l" = default ]; then cp /tmp/default.so jpeg-encoder-and-decoder_b200/libjpgb200.so; else cp profiles/ab/libs/$l.so jpeg-encoder-and-decoder_b200/libjpgb200.so; fi
  if [ -n "$TESTS" ]; then python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic_vs_oracle or tma_and_plain or config2" 2>&1 | tail -1; fi
  python bench.py --frames 64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-worst-case > gpurun_out/${tag}_$l.json 2> gpurun_out/${tag}_$l.err
  python - <<PY
import json
d=[json.loads(x) for x in open("gpurun_out/${tag}_$l.json") if x.strip().startswith("{")][-1]
print("$l", round(d["value"]), {k:round(v/64*1000,1) for k,v in d["roofline"]["kernel_ms_per_step_untimed_pass"].items() if v})
PY
done
cp /tmp/default.so jpeg-encoder-and-decoder_b200/libjpgb200.so
