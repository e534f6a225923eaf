#!/usr/bin/env python3
"""Prints DESIGN.md section 7's table from the bench lines of a measurement set: python tools/design_table.py profiles/r01_final4"""
import json, os, sys
d = sys.argv[1]
rows = [('cfg1_scene0_512', 'cfg1 `scene0` 512²'), ('cfg2_scene1_1080p', '**cfg2 `scene1` 1080p** (default bench)'),
        ('cfg3_scene9_mandelbulb_1080p', 'cfg3 `scene9` mandelbulb 1080p'), ('cfg4a_scene10_menger_1080p_pl32', 'cfg4a `scene10` menger 1080p, pathLength 32'),
        ('cfg4b_scene8_terrain_1080p_pl32', 'cfg4b `scene8` terrain 1080p, pathLength 32'), ('cfg5_scene10_4k', 'cfg5 `scene10` 4K'),
        ('bvh_spheres169_1080p', '169 spheres 1080p (`scenes_synthetic/`)'), ('bvh_mixed74_1080p', '74 mixed primitives 1080p (`scenes_synthetic/`)')]
print('| workload | driver | Gsamples/s | ms/step (64 spp) | e2e Gsamples/s | algorithmic flops/sample | TFLOP/s (algorithmic) | fraction of the 74.4 TFLOP/s FP32-issue roofline | CPU oracle Msamples/s | e2e / CPU |')
print('|---|---|---|---|---|---|---|---|---|---|')
for key, label in rows:
    f = os.path.join(d, 'bench_%s.json' % key)
    if not os.path.exists(f):
        continue
    b = json.load(open(f))
    cfg, rf, cpu = b['config'], b['roofline'], b.get('cpu_baseline') or {}
    drv = 'v1' if (not cfg.get('has_sdf') and key in ('cfg1_scene0_512', 'cfg2_scene1_1080p', 'bvh_spheres169_1080p')) else 'v2s'
    if cfg.get('closest_hit') == 'bvh':
        drv += ' + BVH'
    cv = cpu.get('value')
    print('| %s | %s | %.2f | %.2f | %.2f | %d | %.1f | %.3f | %s | %s |' % (
        label, drv, b['value'] / 1e9, b['ms_per_step'], b['e2e']['value'] / 1e9, round(rf['flops_per_sample_algorithmic']),
        rf['achieved'], rf['frac'], ('%.2f (%d cores)' % (cv / 1e6, cpu.get('cores', 0))) if cv else 'n/a',
        ('%d×' % round(b['e2e']['value'] / cv)) if cv else 'n/a'))
