#!/usr/bin/env python3
"""Second half of BASELINE.json's metric: wall time to reach a fixed relative RMSE against a 65 536-spp reference.

The reference image is rendered in STRICT mode (--ref-mode; round 1 rendered it with the same fast kernel it was then
compared with, which made the curves blind to any bias of the fast build), at --scale of the workload's resolution
when the full frame would take too long in strict mode.
For each workload: the reference image is rendered with the DISJOINT sample indices [2^20, 2^20 + ref_spp) (so test
renders are statistically independent of it), then test renders with indices 0..spp-1 for
spp = 1, 2, 4, ...; relRMSE = sqrt(mean((I-R)^2 / (R^2 + eps))), eps = (0.01 mean R)^2, on the XYZ buffer (computed
on the GPU with torch).  Device time per render comes from CUDA events on the library's stream.
Prints one JSON line per workload: the (spp, seconds, relRMSE) curve and the time to each threshold (log-log
interpolated between the bracketing points)."""
import argparse, json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pathtracer_b200 as pt
from bench import WORKLOADS


def rel_rmse(img, ref):
    a, r = img[..., :3].double(), ref[..., :3].double()
    eps = (0.01 * r.mean()) ** 2
    return float(torch.sqrt(torch.mean((a - r) ** 2 / (r ** 2 + eps))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('workloads', nargs='*', default=['cfg1_scene0_512'])
    ap.add_argument('--ref-spp', type=int, default=65536)
    ap.add_argument('--max-spp', type=int, default=4096)
    ap.add_argument('--mode', default='fast')
    ap.add_argument('--ref-mode', default='strict')
    ap.add_argument('--scale', type=int, default=1, help='render at 1/scale of the resolution in x and y')
    ap.add_argument('--thresholds', default='0.2,0.1,0.05,0.02')
    args = ap.parse_args()
    thresholds = [float(t) for t in args.thresholds.split(',')]
    for wl in args.workloads:
        scene, W, H, _, pl, _, _ = WORKLOADS[wl]
        W, H = W // args.scale, H // args.scale
        sc = pt.Scene.load(os.path.join(ROOT, 'scenes', scene + '.json'))
        modes = {'fast': pt.MODE_FAST, 'strict': pt.MODE_STRICT}
        r = pt.Renderer(mode=modes[args.mode], jit=2)
        r.set_scene(sc.pack_ubo(), sc.sdf_sources)
        rr = pt.Renderer(mode=modes[args.ref_mode], jit=2)
        rr.set_scene(sc.pack_ubo(), sc.sdf_sources)
        p = sc.pack_params(1, W, H, 16, pl)
        img = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
        r.bind_image(img)
        rr.bind_image(img)

        def render(first, spp, r=r):
            img.zero_()
            r.sync(); r.kernel_time()
            s = first
            while s < first + spp:
                n = min(64, first + spp - s)
                r.dispatch_sum(p, s, n)
                s += n
            r.finalize(p, spp)
            ms, _ = r.kernel_time()
            return img.clone(), ms * 1e-3

        ref, ref_s = render(1 << 20, args.ref_spp, rr)
        curve = []
        spp = 1
        while spp <= args.max_spp:
            im, secs = render(0, spp)
            curve.append((spp, secs, rel_rmse(im, ref)))
            spp *= 2
        out = {'workload': wl, 'scene': scene, 'width': W, 'height': H, 'path_length': pl, 'mode': args.mode, 'ref_mode': args.ref_mode,
               'reference': {'spp': args.ref_spp, 'first_sample': 1 << 20, 'seconds': ref_s},
               'curve': [{'spp': s, 'seconds': t, 'rel_rmse': e} for s, t, e in curve], 'time_to_rel_rmse': {}}
        for th in thresholds:
            tt = None
            for (s0, t0, e0), (s1, t1, e1) in zip(curve, curve[1:]):
                if e0 > th >= e1:
                    f = (math.log(e0) - math.log(th)) / (math.log(e0) - math.log(e1))
                    tt = math.exp(math.log(t0) + f * (math.log(t1) - math.log(t0)))
                    break
            if curve and curve[0][2] <= th:
                tt = curve[0][1]
            out['time_to_rel_rmse'][str(th)] = tt
        print(json.dumps(out))
        r.close()
        rr.close()


if __name__ == '__main__':
    main()
