#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 50 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('builtin value', round(d['value']), 'us/frame', round(d['ms_per_step']/16*1000,1))"
python bench.py --steps 10 --scene cornell --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cornell value', round(d['value']), 'us/frame', round(d['ms_per_step']/16*1000,1))"
for n in 20000 500000; do python bench.py --steps 5 --scene mesh --mesh-tris $n --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('mesh $n: value', round(d['value']), 'us/frame', round(d['ms_per_step']/16*1000,1), 'Mrays/s', round(d['mrays_per_s']))"; done
