#!/bin/bash
# usage (on the GPU box): bash scripts/sweep_variants.sh variants/*.so  -> one line per variant
for v in "$@"; do
  EMRIFD_LIB=$PWD/$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /tmp/out.json 2> /tmp/err.log
  if [ -s /tmp/out.json ]; then
    python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/out.json"))
print(sys.argv[1], "walkers/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "kernel_ms", round(d["roofline"]["kernel_ms"], 3), "e2e", round(d["e2e"]["value"], 1))
PY
  else
    echo "$v FAILED: $(tail -2 /tmp/err.log | tr '\n' ' ')"
  fi
done
