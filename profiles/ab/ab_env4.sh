# usage: ab_env4.sh tag "ENV=..." "ENV=..."   ("-" = no env)
tag=$1; shift
i=0
for e in "$@"; do
  i=$((i+1))
  if [ "$e" = "-" ]; then ee=""; else ee="$e"; fi
  env $ee python bench.py --frames 64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-worst-case > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - <<PY
import json
d=[json.loads(x) for x in open("gpurun_out/${tag}_$i.json") if x.strip().startswith("{")][-1]
print("$e", round(d["value"]), {k:round(v/64*1000,1) for k,v in d["roofline"]["kernel_ms_per_step_untimed_pass"].items() if v})
PY
done
