"""The N>1 path.  CPU: the slicing plan and a world_size-2 gloo run of the reduce + finalize arithmetic, with the CPU
oracle standing in for the kernel (test infrastructure only).  GPU (-m gpu, needs >= 2 devices): the real thing."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_path
from pathtracer_b200 import multi


def test_slices_tile_the_range():
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 64, 1024, 16384):
            edges = [multi.sample_slice(r, world, total, first_sample=5) for r in range(world)]
            assert edges[0][0] == 5 and edges[-1][1] == 5 + total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    assert multi.chunks(3, 20, 8) == [(3, 8), (11, 8), (19, 1)]
    assert multi.chunks(4, 4, 8) == []
    with pytest.raises(ValueError):
        multi.sample_slice(2, 2, 10)


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from oracle import oracle, pack
from pathtracer_b200 import multi
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
o, scene = oracle.from_scene_file(%(scene)r, threads=2)
W, H, total, spf = 32, 24, 10, 4
p = pack.pack_params(scene, 1, W, H, spf, 5)
img = np.zeros((H, W, 4), dtype=np.float32)
b, e = multi.sample_slice(rank, world, total)
for s, n in multi.chunks(b, e, spf):
    o.dispatch_sum(p, s, n, img)           # the CPU oracle stands in for Renderer.dispatch_sum
t = torch.from_numpy(img)
dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
if rank == 0:
    np.save(%(out)r, t.numpy())
dist.barrier()
dist.destroy_process_group()
'''


def test_world_size_2_gloo_reduce_matches_single_process(tmp_path):
    from oracle import oracle, pack
    out = str(tmp_path / 'sum.npy')
    port = 29500 + (os.getpid() % 2000)
    code = WORKER % {'root': ROOT, 'port': port, 'scene': scene_path('scene1'), 'out': out}
    script = tmp_path / 'worker.py'
    script.write_text(code)
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for pr in procs:
        assert pr.wait(timeout=300) == 0
    got = np.load(out)
    o, scene = oracle.from_scene_file(scene_path('scene1'))
    W, H, total = 32, 24, 10
    p = pack.pack_params(scene, 1, W, H, 4, 5)
    ref = np.zeros((H, W, 4), dtype=np.float32)
    o.dispatch_sum(p, 0, total, ref)
    # same samples, different fp32 summation order: tolerance, not bits
    assert np.allclose(got[..., :3], ref[..., :3], rtol=1e-5, atol=1e-7)
    # finalize: sum / total * apertureSize^2 * ISO equals the running-mean render of the same samples
    expo = np.float32(p['apertureSize']) ** 2 * np.float32(int(p['ISO']))
    fin = got[..., :3] / np.float32(total) * expo
    mean = o.render(pack.pack_params(scene, 1, W, H, 5, 5), total, 5)
    assert np.allclose(fin, mean[..., :3], rtol=2e-5, atol=1e-8)


@pytest.mark.gpu
def test_two_gpu_split_matches_one_gpu(ptlib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'w.py'
    out = str(tmp_path / 'img.npy')
    script.write_text(r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import pathtracer_b200 as pt
from pathtracer_b200 import multi
dist.init_process_group('nccl')
lr = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(lr)
sc = pt.Scene.load(%r); ubo = sc.pack_ubo(); p = sc.pack_params(1, 160, 90, 8, 5)
r = pt.Renderer(device=lr, mode=pt.MODE_STRICT); r.set_scene(ubo, sc.sdf_sources)
img = torch.zeros((90, 160, 4), dtype=torch.float32, device='cuda')
multi.render_split(r, p, 32, 8, img, dist)
if dist.get_rank() == 0: np.save(%r, img.cpu().numpy())
dist.barrier(); dist.destroy_process_group()
''' % (ROOT, scene_path('scene10'), out))
    subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                    '--master-port', '29611', str(script)], check=True, timeout=600)
    got = np.load(out)
    import torch
    sc = ptlib.Scene.load(scene_path('scene10'))
    r = ptlib.Renderer(device=0, mode=ptlib.MODE_STRICT)
    r.set_scene(sc.pack_ubo(), sc.sdf_sources)
    img = torch.zeros((90, 160, 4), dtype=torch.float32, device='cuda')
    multi.render_split(r, sc.pack_params(1, 160, 90, 8, 5), 32, 8, img, None)
    one = img.cpu().numpy()
    assert np.allclose(got, one, rtol=1e-5, atol=1e-8)
    assert (got[..., 3] == 1.0).all()
