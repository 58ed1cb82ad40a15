#!/bin/bash
# round-2 re-entry check on one B200: the new integrator_Hart tests first, then the whole GPU suite with
# its slowest tests listed, smoke(), and the default bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "hart or spirv_pins or unsupported" 2>&1 | tail -5
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2k_tests.log 2>&1
grep -E "passed|failed|error|real" gpurun_out/r2k_tests.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_r2k_n1.json 2> gpurun_out/bench_r2k_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_r2k_n1.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "frac", round(d["roofline"]["frac"], 3),
      "c3", round(d["c3"]["value"]), d["c3"]["parity_ok"], "c4", round(d["c4"]["value"]), d["c4"]["parity_ok"], d["clocks"], d["cpu_baseline"]["value"])
PY
