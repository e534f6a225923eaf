#!/usr/bin/env python3
"""Compile the scene-specialised kernel on the CPU box (NVRTC) and report registers + static code size.
usage: python tools/jit_size.py scene1 [fast|strict] [key=value ...]   (options as for pt_set_option, e.g. sched=8)"""
import glob, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
d = tempfile.mkdtemp()
os.environ['PT_JIT_DUMP'] = d   # debugging aid of pt_jit.cpp: keeps the generated source and the cubin
import pathtracer_b200 as pt
name = sys.argv[1]
mode = 0 if (len(sys.argv) > 2 and sys.argv[2] == 'strict') else 1
opts = {k: int(v) for k, v in (a.split('=', 1) for a in sys.argv[2:] if '=' in a)}
path = name if os.path.exists(name) else os.path.join(ROOT, 'scenes/%s.json' % name)
sc = pt.Scene.load(path)
log = pt.kernel_compile_check(sc.pack_ubo(), sc.sdf_sources, mode, True, opts)
print(' | '.join(l.strip() for l in log.split('\n') if 'Used' in l or 'spill' in l))
cub = glob.glob(d + '/*.cubin')[0]
subprocess.run([sys.executable, os.path.join(ROOT, 'tools/sass_size.py'), cub, 'pt_render_jit', os.path.join(ROOT, 'pathtracer_b200/csrc/pt_kernel.cuh')])
