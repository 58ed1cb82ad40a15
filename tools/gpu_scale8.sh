#!/bin/bash
# 8-GPU box: image check (p2p + nccl assembly vs 1-GPU image), strong scaling at 1080p for N = 2, 4, 8, C4 (4K) at N = 1 and 8
mkdir -p gpurun_out
TAG=${1:-r01}
N=8
sed -n '/^cat > \/tmp\/mg_check.py/,/^PY$/p' tools/gpu_multi.sh > /tmp/mk.sh; N=8 bash /tmp/mk.sh
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 /tmp/mg_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -4
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then timeout 200 python bench.py --steps 60 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$n.json 2>gpurun_out/scale_${TAG}_n$n.err;
  else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 60 > gpurun_out/scale_${TAG}_n$n.json 2>gpurun_out/scale_${TAG}_n$n.err; fi
  grep -iE "error|Traceback" gpurun_out/scale_${TAG}_n$n.err | head -3
  python - <<PY
import json
l=[x for x in open('gpurun_out/scale_${TAG}_n$n.json') if x.startswith('{')]
d=json.loads(l[-1]); print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'us/frame', round(d['ms_per_step']/d['config']['frames_per_step']*1000,1), '|', d['config']['gather'][:30], d['clocks']['sm_mhz'])
PY
done
for n in 1 8; do
  if [ $n -eq 1 ]; then timeout 200 python bench.py --steps 20 --width 3840 --height 2160 --no-cpu-baseline > gpurun_out/c4_${TAG}_n$n.json 2>gpurun_out/c4_${TAG}_n$n.err;
  else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --width 3840 --height 2160 > gpurun_out/c4_${TAG}_n$n.json 2>gpurun_out/c4_${TAG}_n$n.err; fi
  python - <<PY
import json
l=[x for x in open('gpurun_out/c4_${TAG}_n$n.json') if x.startswith('{')]
d=json.loads(l[-1]); print('C4 4K N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'us/frame', round(d['ms_per_step']/d['config']['frames_per_step']*1000,1))
PY
done
