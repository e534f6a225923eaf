#!/bin/bash
BENCH_DEBUG=1 python bench.py --no-cpu-baseline --steps 8 --workload cfg5_scene10_4k > gpurun_out/x_cfg5.json 2>gpurun_out/x.err
cat gpurun_out/x.err | tail -12
python - <<'PY' 2>&1 | tail -12
import time, torch, numpy as np, sys
sys.path.insert(0,'.')
import pathtracer_b200 as pt
sc = pt.Scene.load('scenes/scene10.json'); ubo = sc.pack_ubo(); p = sc.pack_params(1,3840,2160,16,5)
r = pt.Renderer(mode=pt.MODE_FAST, jit=2); r.set_scene(ubo, sc.sdf_sources); r.resize(3840,2160)
h = [torch.empty((2160,3840,4),dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
r.dispatch(p); r.sync()
t=time.perf_counter(); r.read_xyz(h[0]); print('blocking read %.2f ms'%(1e3*(time.perf_counter()-t)))
t=time.perf_counter(); r.read_xyz_async(h[0]); r.read_wait(); print('async read alone %.2f ms (first)'%(1e3*(time.perf_counter()-t)))
t=time.perf_counter(); r.read_xyz_async(h[1]); r.read_wait(); print('async read alone %.2f ms'%(1e3*(time.perf_counter()-t)))
t=time.perf_counter(); r.dispatch(p); r.sync(); print('dispatch alone %.2f ms'%(1e3*(time.perf_counter()-t)))
t=time.perf_counter(); r.dispatch(p); r.read_xyz_async(h[0]); r.dispatch(p); r.sync(); r.read_wait(); print('dispatch+async read+dispatch %.2f ms'%(1e3*(time.perf_counter()-t)))
t=time.perf_counter(); r.set_scene(ubo, sc.sdf_sources); print('set_scene %.2f ms'%(1e3*(time.perf_counter()-t)))
PY
