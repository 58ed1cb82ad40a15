#!/bin/bash
# round 2, second GPU pass: all parity tests, the reworked bench (default line), timelines of the batched kernel
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
tail -c 4500 gpurun_out/bench_${TAG}_n1.json; tail -3 gpurun_out/bench_${TAG}_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>&1
cut -c1-400 gpurun_out/bench_${TAG}_ref.json
timeout 300 python bench.py --graph on --no-cpu-baseline --no-parity --no-c4 > gpurun_out/bench_${TAG}_graph.json 2>gpurun_out/bench_${TAG}_graph.err
python -c "
import json;d=json.load(open('gpurun_out/bench_${TAG}_graph.json'));print('graph on: value',round(d['value']),'e2e',round(d['e2e']['value']))"
timeout 200 python tools/timeline.py --batch 22 > gpurun_out/timeline_${TAG}_builtin_b22.md 2>&1
timeout 200 python tools/timeline.py --batch 8 --nranks 8 > gpurun_out/timeline_${TAG}_builtin_b8_n8.md 2>&1
timeout 200 python tools/timeline.py --batch 22 --nranks 8 > gpurun_out/timeline_${TAG}_builtin_b22_n8.md 2>&1
timeout 200 python tools/timeline.py --batch 15 --scene cornell > gpurun_out/timeline_${TAG}_cornell_b15.md 2>&1
cat gpurun_out/timeline_${TAG}_builtin_b22.md gpurun_out/timeline_${TAG}_builtin_b22_n8.md gpurun_out/timeline_${TAG}_cornell_b15.md
