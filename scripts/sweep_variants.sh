#!/bin/bash
# usage (on the GPU box): bash scripts/sweep_variants.sh variants/*.so  -> one line per variant
for v in "$@"; do
  EMRIFD_LIB=$PWD/$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /tmp/out.json 2> /tmp/err.log
  if [ -s /tmp/out.json ]; then
    python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/out.json"))
c1, c4, ep = d.get("cfg1", {}), d.get("cfg4_binsharded", {}), d.get("e2e_from_parameters", {})
print(sys.argv[1], "walkers/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "kernel_ms", round(d["roofline"]["kernel_ms"], 3),
      "e2e", round(d["e2e"]["value"], 1), "| cfg1 kernel_ms", round(c1.get("kernel_ms") or 0, 4), "sum", round(c1.get("mode_sum_ms") or 0, 4), "hbm_frac", round(c1.get("hbm_frac") or 0, 3),
      "| cfg4 ms", round(c4.get("ms_per_likelihood") or 0, 3), "| from_params", round(ep.get("value") or 0, 1), ep.get("error") or "")
PY
  else
    echo "$v FAILED: $(tail -2 /tmp/err.log | tr '\n' ' ')"
  fi
done
