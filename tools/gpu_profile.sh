#!/bin/bash
# bench + ncu launch list + ncu full capture of the top kernels + sanitizer on smoke()
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
tail -c 3000 gpurun_out/bench_${TAG}_n1.json
tail -5 gpurun_out/bench_${TAG}_n1.err
python bench.py --steps 10 --warmup 3 --pose pinned --no-cpu-baseline > gpurun_out/bench_${TAG}_pinned.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --frames 4 --no-cpu-baseline --graph off > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 8 -c 2 \
    -f -o gpurun_out/prof_frame_${TAG} python bench.py --steps 1 --warmup 3 --frames 4 --no-cpu-baseline --graph off > gpurun_out/ncu_frame.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${TAG}.log 2>&1
tail -5 gpurun_out/sanitizer_${TAG}.log
ls -la gpurun_out
