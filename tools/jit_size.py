#!/usr/bin/env python3
"""Compile the scene-specialised kernel on the CPU box (NVRTC) and report registers + static code size.
usage: PT_SCHED=0/1 ... python tools/jit_size.py scene1 [fast|strict]"""
import ctypes as C, glob, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
d = tempfile.mkdtemp()
os.environ['PT_JIT_DUMP'] = d
import pathtracer_b200 as pt
from pathtracer_b200 import api
name = sys.argv[1]; mode = 1 if (len(sys.argv) < 3 or sys.argv[2] == 'fast') else 0
sc = pt.Scene.load(os.path.join(ROOT, 'scenes/%s.json' % name)); ubo = sc.pack_ubo()
L = pt.lib(); L.pt_kernel_compile_check.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]
src = sc.sdf_sources
rc = L.pt_kernel_compile_check(ubo.ctypes.data_as(C.c_void_p), api._c_strings(src), len(src), mode, 1)
log = L.pt_last_error(None).decode()
if rc: print(log); sys.exit(1)
print(' | '.join(l.strip() for l in log.split('\n') if 'Used' in l or 'spill' in l))
cub = glob.glob(d + '/*.cubin')[0]
subprocess.run([sys.executable, os.path.join(ROOT, 'tools/sass_size.py'), cub, 'pt_render_jit', os.path.join(ROOT, 'pathtracer_b200/csrc/pt_kernel.cuh')])
