#!/bin/bash
# multi-GPU check: N ranks, NCCL gather; compares the assembled image with a 1-GPU render
N=${1:-2}
mkdir -p gpurun_out
cat > /tmp/mg_check.py <<PY
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
import rvpt_b200 as rv
from rvpt_b200.distributed import FrameGather, PeerOutput
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
W, H = 1920, 1080
s = rv.builtin_scene(); nodes, perm = rv.build_bvh(s.triangles); tris = s.triangles[perm]
cam = rv.camera_data(aspect=W / H)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
want = None
if rank == 0:
    ref = rv.Engine(W, H, device=lr); ref.set_stream(st.cuda_stream)
    ref.upload_scene(tris, s.materials, nodes)
    for f in range(5): ref.render_frame(rv.default_settings(frame=f), cam)
    want = ref.read_output_rgba8()
for mode in ('nccl', 'p2p'):
    eng = rv.Engine(W, H, device=lr, rank=rank, nranks=world); eng.set_stream(st.cuda_stream)
    eng.upload_scene(tris, s.materials, nodes)
    if mode == 'nccl':
        fg = FrameGather(eng, dist, torch, torch.device('cuda', lr))
        for f in range(5):
            fg.begin_frame(); eng.render_frame(rv.default_settings(frame=f), cam); fg.end_frame()
        img = fg.image()
    else:
        po = PeerOutput(eng, dist, torch, torch.device('cuda', lr))
        for f in range(5): eng.render_frame(rv.default_settings(frame=f), cam)
        img = po.image()
    if rank == 0:
        print(mode, 'multi-GPU image equals 1-GPU image:', np.array_equal(img, want), 'ranks', world, flush=True)
        assert np.array_equal(img, want)
    dist.barrier()
dist.destroy_process_group()
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 /tmp/mg_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6
run() { # n extra-args tag
  n=$1; tag=$3
  if [ $n -eq 1 ]; then timeout 200 python bench.py --steps 30 --no-cpu-baseline $2 > gpurun_out/scale_$tag.json 2>gpurun_out/scale_$tag.err;
  else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 30 $2 > gpurun_out/scale_$tag.json 2>gpurun_out/scale_$tag.err; fi
  grep -iE "error|Traceback" gpurun_out/scale_$tag.err | head -3
  python - <<PY
import json
l=[x for x in open('gpurun_out/scale_$tag.json') if x.startswith('{')]
d=json.loads(l[-1]); print('$tag: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'us/frame', round(d['ms_per_step']/d['config']['frames_per_step']*1000,1), '|', d['config']['gather'][:40], '|', d['config']['launch'][:40])
PY
}
run 1 "" n1
run $N "--gather none" n${N}_nogather
run $N "--gather nccl" n${N}_nccl
run $N "--gather p2p --graph off" n${N}_p2p_nograph
run $N "" n${N}_p2p
