#!/bin/bash
# Round-2 artefacts on one B200: bench lines, ncu launch list, ncu full capture of the batched frame kernel,
# phase timelines, bounce sweep, compute-sanitizer. Raw files land in gpurun_out/; tools/summarize_profile.py r02
# turns them into profiles/.
mkdir -p gpurun_out
TAG=${1:-r02}
[ -n "$SKIP_TESTS" ] || timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null
timeout 300 python bench.py --pose pinned --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_pinned.json 2>/dev/null
timeout 600 python bench.py --scene cornell --steps 5 --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_cornell.json 2>/dev/null
timeout 300 python bench.py --frame-by-frame --no-cpu-baseline --no-c4 --no-parity > gpurun_out/bench_${TAG}_frame_by_frame.json 2>/dev/null
timeout 600 python bench.py --scene mesh --mesh-tris 500000 --frames 16 --steps 3 --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_mesh500k.json 2>/dev/null
timeout 600 python bench.py --scene tridel --frames 16 --steps 3 --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_tridel.json 2>/dev/null
python - "$TAG" <<'PY'
import json, sys
for n in ("n1", "pinned", "cornell", "frame_by_frame", "mesh500k", "tridel"):
    try:
        d = json.loads([l for l in open("gpurun_out/bench_%s_%s.json" % (sys.argv[1], n)) if l.startswith("{")][-1])
        r = d["roofline"]
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "frac", round(r["frac"], 3),
              "ms/launch", round(r["ms_per_launch"], 3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 --no-parity > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 4 -c 1 -f -o gpurun_out/prof_frame_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 --no-parity > gpurun_out/ncu_frame.log 2>&1
tail -2 gpurun_out/ncu_frame.log
timeout 200 python tools/timeline.py --batch 64 > gpurun_out/timeline_${TAG}_builtin_b64.md 2>&1
timeout 200 python tools/timeline.py --batch 64 --nranks 8 > gpurun_out/timeline_${TAG}_builtin_b64_rank0of8.md 2>&1
timeout 200 python tools/timeline.py --batch 64 --pose pinned > gpurun_out/timeline_${TAG}_pinned_b64.md 2>&1
timeout 200 python tools/timeline.py --batch 32 --scene cornell > gpurun_out/timeline_${TAG}_cornell_b32.md 2>&1
timeout 200 python tools/timeline.py > gpurun_out/timeline_${TAG}_builtin_frame_by_frame.md 2>&1
timeout 600 python tools/bounce_sweep.py ${TAG} > gpurun_out/bounce_sweep_${TAG}.log 2>&1; cp profiles/${TAG}_bounce_sweep.md gpurun_out/ 2>/dev/null
bash tools/gpu_sanitize.sh ${TAG}
ls gpurun_out | grep ${TAG} | head -50
