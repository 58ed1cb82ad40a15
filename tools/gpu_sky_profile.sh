#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
cat > /tmp/sky.py <<PY
import sys; sys.path.insert(0,'.')
import numpy as np, rvpt_b200 as rv
s=rv.builtin_scene(); nodes,perm=rv.build_bvh(s.triangles); tris=s.triangles[perm]
W,H=1920,1080
cam=rv.camera_data(translation=(0,50.0,0),aspect=W/H)
e=rv.Engine(W,H); e.upload_scene(tris,s.materials,nodes)
for f in range(12): e.render_frame(rv.default_settings(max_bounces=1,frame=f),cam)
e.sync()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 8 -c 1 -f -o gpurun_out/prof_sky python /tmp/sky.py > gpurun_out/ncu_sky.log 2>&1
tail -2 gpurun_out/ncu_sky.log
