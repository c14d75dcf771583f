#!/bin/bash
# Dev diagnostic: does the NVML polling thread of bench.py's ClockSampler slow the timed region down?
cd "$(dirname "$0")/.."
for p in 1 10 0; do
  python bench.py --clock-period-ms $p --no-cpu-baseline --no-e2e --no-render 2>/dev/null > /tmp/cp.json
  python - "$p" <<'PY'
import json, sys
d = json.load(open('/tmp/cp.json'))
print("period_ms", sys.argv[1], "us/step", round(d["ms_per_step"] * 1e3, 2), "kernel_us", round(d["roofline"]["kernel_ms"] * 1e3, 2), d["clocks"])
PY
done
python bench.py --clock-period-ms 10 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-render 2>/dev/null > /tmp/cp.json
python - <<'PY'
import json
d = json.load(open('/tmp/cp.json'))
print("steps 20, period 10 ms: us/step", round(d["ms_per_step"] * 1e3, 2), d["clocks"])
PY
