#!/usr/bin/env python3
"""bench.py -- spectral path samples/s of libpt_cuda on BASELINE.json's configs (driver contract: see DESIGN.md).

  python bench.py --gpus 1 --steps K --warmup W            our arm, one B200
  torchrun ... bench.py --gpus N --steps K --warmup W      our arm, one rank per GPU (sample-split, one NCCL reduce)
  python bench.py --impl reference ...                     the reference algorithm on the host cores (CPU oracle)

A "step" is one dispatch of the hot path: --spf samples per pixel over the whole frame of the workload
(default: BASELINE config 2, scenes/scene1.json at 1920x1080, pathLength 5, camera shot 1; 16 steps x 64 = 1024 spp).
1 sample = one Scene() call (shader.comp:1446-1490): one 4-wavelength hero bundle through the camera lens and the
whole path.  `value` is device time (CUDA events on the library's stream, inputs resident in HBM, L2 flushed between
steps); `e2e` is the same metric through the C ABI with host buffers (scene block upload + dispatch + read-back of
the XYZ image into pinned host memory every step).  Data is synthetic in the sense of the contract: the reference's
own shipped scene files, no external assets.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene, width, height, total spp of the config, pathLength, C_sdf, C_mat)   
    'cfg1_scene0_512': ('scene0', 512, 512, 64, 5, 0, 0),
    'cfg2_scene1_1080p': ('scene1', 1920, 1080, 1024, 5, 0, 0),
    'cfg3_scene9_mandelbulb_1080p': ('scene9', 1920, 1080, 1024, 5, 270, 11),
    'cfg4a_scene10_menger_1080p_pl32': ('scene10', 1920, 1080, 1024, 32, 214, 0),
    'cfg4b_scene8_terrain_1080p_pl32': ('scene8', 1920, 1080, 1024, 32, 92, 0),
    'cfg5_scene10_4k': ('scene10', 3840, 2160, 16384, 5, 214, 0),
    # beyond the shipped scenes (section 8f-3): primitive counts where the BVH replaces the scan; --bvh-min 0 = scan
    'bvh_spheres169_1080p': ('synthetic/spheres169', 1920, 1080, 1024, 5, 0, 0),
    'bvh_mixed74_1080p': ('synthetic/mixed74', 1920, 1080, 1024, 5, 0, 0),
}
N_SM, FP32_LANES = 148, 128


def scene_file(name):
    """'sceneN' -> the reference's shipped scenes/sceneN.json; 'synthetic/X' -> scenes_synthetic/X.json"""
    if name.startswith('synthetic/'):
        return os.path.join(ROOT, 'scenes_synthetic', name.split('/', 1)[1] + '.json')
    return os.path.join(ROOT, 'scenes', name + '.json')


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner when
    NCCL_DEBUG=VERSION is set on the box), so file descriptor 1 is pointed at stderr for the duration of the run and
    the line goes to a saved copy of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0}, 'fallback'


def flops_per_sample(cnt, scene_counts, c_sdf, c_mat):
    """Algorithmic flops per sample: oracle call counters x the per-call constants of SURVEY.md App. D (v1)."""
    nS, nP, nB, nL, nC, nSdf = scene_counts
    rays = cnt['rays_path'] + cnt['rays_shadow']
    f = 1203.0 * cnt['samples']
    f += rays * 3.0
    f += cnt['sphere'] * 29.0 + cnt['sphere_hit'] * 18.0
    f += cnt['plane'] * 6.0 + cnt['plane_hit'] * 6.0
    f += rays * nB * 24.0 + cnt['box'] * 167.0 + cnt['box_hit'] * 38.0
    f += rays * nL * 22.0 + cnt['lens'] * 347.0 + cnt['slice_hit'] * 33.0
    f += rays * nC * 18.0 + cnt['cyclide'] * 600.0 + cnt['cyclide_3root'] * 60.0 + cnt['cyclide_hit'] * 51.0
    f += cnt['searchsdf'] * 37.0 * nSdf + cnt['st_calls'] * 3.0 + cnt['st_enter'] * 11.0
    f += (cnt['st_iter'] - cnt['st_backstep']) * 25.0 + cnt['st_backstep'] * 13.0
    f += cnt['sdf_eval'] * (c_sdf + 4.0) + cnt['st_hit'] * 73.0 + cnt['sdfmat_eval'] * (2.0 * c_sdf + c_mat)
    f += cnt['bounce'] * 229.0 + cnt['light_visible'] * 103.0 + cnt['emit_hit'] * 70.0
    return f / max(cnt['samples'], 1)


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  Polled through
    NVML in a thread (nvidia_ml_py, every 10 ms: also a 100 ms region gets its samples); if NVML cannot be used the
    sampler falls back to `nvidia-smi -lms 50` (started before the warm-up: it takes a moment to come up and its
    rows arrive through a pipe).  Only samples stamped inside [mark_begin, mark_end] are used."""

    REASONS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))

    def __init__(self, index, uuid=None):
        self.rows = []      # (time, sm_mhz, sm_max_mhz, power_w, [reason names])
        self.proc = None
        self.index = index  # NVML / nvidia-smi index; `uuid` (of the CUDA device) wins when NVML knows it
        self.uuid = uuid
        self.t0 = self.t1 = None
        self.source = None
        self._stop = threading.Event()
        self._thread = None

    # ---- NVML ------------------------------------------------------------------------------------------------
    def _start_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = None
        if self.uuid:
            try:
                h = nv.nvmlDeviceGetHandleByUUID(('GPU-' + str(self.uuid)).encode())
            except Exception:
                h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))  # probe once: raises here, not in the thread
        int(get_reasons(h))

        def poll():
            while not self._stop.is_set():
                try:
                    sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    mask = int(get_reasons(h))
                    try:
                        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                    except Exception:
                        pw = None
                    self.rows.append((time.time(), sm, mx, pw, [n for bit, n in self.REASONS if mask & bit]))
                except Exception:
                    pass
                self._stop.wait(0.01)

        self._thread = threading.Thread(target=poll, daemon=True)
        self._thread.start()
        self.source = 'nvml'

    # ---- nvidia-smi ------------------------------------------------------------------------------------------
    def _start_smi(self):
        q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
                                      '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

        def read():
            for line in self.proc.stdout:
                r = [t.strip() for t in line.split(',')]
                try:
                    reasons = [n for n, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8])
                               if v.lower().startswith('active')]
                    self.rows.append((time.time(), float(r[1]), float(r[2]), float(r[3]), reasons))
                except (ValueError, IndexError):
                    continue

        self._thread = threading.Thread(target=read, daemon=True)
        self._thread.start()
        self.source = 'nvidia-smi'

    def start(self):
        try:
            self._start_nvml()
            return
        except Exception:
            self.source = None
        try:
            self._start_smi()
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.12)  # let the last rows of the window arrive
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        rows = list(self.rows)
        slack = 0.0 if self.source == 'nvml' else 0.1
        inside = [r for r in rows if self.t0 is not None and self.t0 - slack / 2 <= r[0] <= (self.t1 or r[0]) + slack]
        used = inside or rows[-3:]
        if not used:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'source': self.source}
        reasons = sorted({n for r in used for n in r[4]})
        power = [r[3] for r in used if r[3] is not None]
        return {'sm_mhz': float(np.median([r[1] for r in used])), 'sm_max_mhz': float(max(r[2] for r in used)), 'reasons': reasons,
                'samples': len(used), 'samples_in_timed_region': len(inside), 'power_w_max': max(power) if power else None,
                'source': self.source}


def ncu_traffic(wl):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the workload's kernel, from the committed
    `ncu --set full` capture of the same bench command (profiles/r01_final4/ or r01_final/ncu_summary_*.json; the texel
    read-modify-write does not depend on the samples per launch); None if not captured."""
    for sub in ('r01_final4', 'r01_final'):
        path = os.path.join(ROOT, 'profiles', sub, 'ncu_summary_%s.json' % wl.split('_')[0])
        try:
            with open(path) as f:
                d = json.load(f)
            if d.get('workload') == wl:
                return d['traffic_bytes_per_launch']
        except (OSError, ValueError, KeyError):
            pass
    return None


def oracle_for(scene_name, count=False, threads=0):
    from oracle import oracle, pack
    scene = pack.load_scene(scene_file(scene_name))
    return oracle.Oracle(pack.pack_ubo(scene), pack.sdf_sources(scene), count=count, threads=threads), scene


def cpu_rate(scene_name, width, height, path_length, target_s, first_dispatch=1):
    """Times the CPU oracle (all host threads) on a bounded sample of the workload sized for about target_s seconds:
    a strided subset of the rows of the full-resolution frame (or, when one whole frame is too quick, several
    samples per pixel of the whole frame).  Returns (samples/s, threads, description)."""
    from oracle import oracle, pack
    o, scene = oracle_for(scene_name)
    threads = oracle.lib().oracle_max_threads()
    img = np.zeros((height, width, 4), dtype=np.float32)
    p = pack.pack_params(scene, 1, width, height, 1, path_length, dispatch=first_dispatch)
    step = max(height // max(2 * threads, 8), 1)      # probe: a few rows per thread
    t0 = time.perf_counter()
    o.dispatch(p, img, 0, step)
    dt = max(time.perf_counter() - t0, 1e-6)
    rate = len(range(0, height, step)) * width / dt
    want = rate * target_s                              # samples that fit the budget
    spf = int(min(max(want // (width * height), 1), 64))
    rows = int(min(max(want // (width * spf), 2 * threads), height))
    step = max(height // rows, 1)
    p = pack.pack_params(scene, 1, width, height, spf, path_length, dispatch=first_dispatch)
    img[:] = 0
    t0 = time.perf_counter()
    o.dispatch(p, img, 0, step)
    dt = time.perf_counter() - t0
    rows = len(range(0, height, step))
    return rows * width * spf / dt, threads, '%d of %d rows (every %d-th) of %s at %dx%d, %d spp, pathLength %d: %.2f s' % (
        rows, height, step, scene_name, width, height, spf, path_length, dt)


def counters_for(scene_name, width, height, path_length, rows=24):
    from oracle import pack
    o, scene = oracle_for(scene_name, count=True)
    p = pack.pack_params(scene, 1, width, height, 1, path_length)
    img = np.zeros((height, width, 4), dtype=np.float32)
    cnt = o.dispatch(p, img, 0, max(height // rows, 1))
    counts = tuple(int(v) for v in o.ubo[:6])
    return cnt, counts


def run_reference(args, wl):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself (Vulkan + glslang + GLFW)
    cannot be built or run in this image, so this is the CPU oracle port, all host threads."""
    scene_name, W, H, spp_cfg, pl, c_sdf, c_mat = WORKLOADS[wl]
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    total = args.steps + args.warmup
    per_step = min(max(90.0 / max(total, 1), 0.05), 5.0)
    rates, desc, threads = [], '', 1
    t_all = time.perf_counter()
    for i in range(total):
        rate, threads, desc = cpu_rate(scene_name, W, H, pl, per_step, first_dispatch=1 + i)
        if i >= args.warmup:
            rates.append(rate)
    wall = time.perf_counter() - t_all
    value = float(len(rates) / sum(1.0 / r for r in rates)) if rates else 0.0
    line = {
        'impl': 'reference', 'metric': 'spectral path samples/sec', 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(total, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic (reference scene files shipped in scenes/)',
        'config': {'workload': wl, 'scene': os.path.relpath(scene_file(scene_name), ROOT), 'width': W, 'height': H, 'path_length': pl,
                   'spp_of_config': spp_cfg, 'note': 'CPU oracle port of shader.comp; each step = a strided-row sample of the frame'},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': threads, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2_scene1_1080p', choices=sorted(WORKLOADS))
    ap.add_argument('--spf', type=int, default=64, help='samples per pixel per step (one dispatch; the sample-stealing driver pools 32 x spf items per warp)')
    ap.add_argument('--mode', default='fast', choices=['fast', 'strict'])
    ap.add_argument('--jit', type=int, default=2, help='0 static kernels, 1 NVRTC for SDF scenes only, 2 NVRTC scene-specialised')
    ap.add_argument('--pipeline', default='megakernel', choices=['megakernel', 'wavefront'])
    ap.add_argument('--bvh-min', type=int, default=None, help='bounded primitives from which the BVH replaces the scan (0: never)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-flush', action='store_true')
    args = ap.parse_args()
    quiet_stdout()
    wl = args.workload
    if args.impl == 'reference':
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    import pathtracer_b200 as pt

    scene_name, W, H, spp_cfg, pl, c_sdf, c_mat = WORKLOADS[wl]
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; libpt_cuda has no CPU fallback (use --impl reference for the CPU oracle)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    sc = pt.Scene.load(scene_file(scene_name))
    ubo = sc.pack_ubo()
    params = sc.pack_params(1, W, H, args.spf, pl)
    r = pt.Renderer(device=local_rank, mode=pt.MODE_FAST if args.mode == 'fast' else pt.MODE_STRICT, jit=args.jit,
                    pipeline=pt.PIPE_WAVEFRONT if args.pipeline == 'wavefront' else pt.PIPE_MEGAKERNEL)
    if args.bvh_min is not None:
        r.set_bvh(args.bvh_min)
    t0 = time.perf_counter()
    r.set_scene(ubo, sc.sdf_sources)
    compile_s = time.perf_counter() - t0
    bvh_active = r.bvh_active
    image = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    r.bind_image(image)
    ext = torch.cuda.ExternalStream(r.stream)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    K, Wm, spf = args.steps, args.warmup, args.spf
    sum_mode = world > 1
    base = rank * (K + Wm) * spf  # this rank's slice of the sample-index range (weak scaling: spp grows with N)

    def step(i):
        if sum_mode:
            r.dispatch_sum(params, base + i * spf, spf)
        else:
            p = params.copy()
            p['frame'] = (i + 1) * spf
            p['currentSamples'] = (i + 1) * spf
            r.dispatch(p)

    def flush_l2():
        if flush is not None:
            with torch.cuda.stream(ext):
                flush.fill_(rank & 0xFF)

    try:
        dev_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    if rank == 0:
        sampler.start()
    for i in range(Wm):
        step(i)
        flush_l2()
    r.sync()
    r.kernel_time()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.mark_begin()
    wall0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for i in range(Wm, Wm + K):
        step(i)
        ms, n = r.kernel_time()  # CUDA events on the library's stream; synchronises
        dev_ms += ms
        launches += n
        flush_l2()
    reduce_ms = 0.0
    if sum_mode:  # the path's one exchange step: NCCL reduce of the per-GPU sum buffers, then finalize on rank 0
        r.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.reduce(image, dst=0, op=dist.ReduceOp.SUM)
        e1.record()
        torch.cuda.synchronize()
        reduce_ms = e0.elapsed_time(e1)
        if rank == 0:
            r.finalize(params, world * (K + Wm) * spf)
            r.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = dev_ms + reduce_ms
    if world > 1:
        t = torch.tensor([total_ms, dev_ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, dev_ms = float(t[0]), float(t[1])
    samples_per_step = W * H * spf
    value = world * samples_per_step * K / (total_ms * 1e-3)

    # ---- e2e: the call a user makes, host buffers in, host buffers out, every step --------------------------------
    hosts = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    hosts_np = [h.numpy() for h in hosts]
    e2e_steps = max(K, 1)                       # every one of the K steps again, end to end
    r.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    r.read_xyz_async(hosts_np[0])               # untimed: creates the copy stream and the snapshot buffer
    r.read_wait()
    dbg = os.environ.get('BENCH_DEBUG')
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r.set_scene(ubo, sc.sdf_sources)       # h2d: the 16 388-byte uniform block (JIT cache hit)
        step(i)                                # the 88-byte push block travels with the launch
        r.read_xyz_async(hosts_np[i & 1])      # d2h: W*H*16 bytes into pinned host memory, overlapping the next step
        if dbg:
            print('e2e step %d issued at %.2f ms' % (i, 1e3 * (time.perf_counter() - t0)), file=sys.stderr)
    r.read_wait()                              # every step's image has landed on the host
    r.sync()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = world * samples_per_step * e2e_steps / e2e_s
    r.kernel_time()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = load_peaks()
    line = {
        'metric': 'spectral path samples/sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic (reference scene files shipped in scenes/, no external assets)',
        'config': {'workload': wl, 'scene': os.path.relpath(scene_file(scene_name), ROOT), 'width': W, 'height': H, 'spf_per_step': spf,
                   'spp_timed': K * spf * world, 'spp_of_config': spp_cfg, 'path_length': pl, 'shot': 1, 'mode': args.mode,
                   'jit': args.jit, 'pipeline': args.pipeline, 'closest_hit': 'bvh' if bvh_active else 'scan', 'l2': 'not flushed' if args.no_flush else 'flushed between steps (256 MiB fill)',
                   'parallelism': 'sample-split x%d + 1 NCCL reduce' % world if world > 1 else 'single GPU',
                   'kernel_compile_s': round(compile_s, 3)},
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': 16388 + 88, 'd2h_bytes_per_step': W * H * 16,
                'steps': e2e_steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'wall_s_timed_region': wall,
        'reduce_ms': reduce_ms,
    }
    # ---- roofline + CPU baseline (rank 0; the oracle is the checker/baseline, never the thing measured above) ------
    try:
        from oracle import oracle as _o
        _o.build()
        cnt, counts = counters_for(scene_name, W, H, pl)
        F = flops_per_sample(cnt, counts, c_sdf, c_mat)
        f_mhz = (clocks or {}).get('sm_mhz') or peaks.get('sm_max_mhz', 1965.0)
        peak_max = N_SM * FP32_LANES * 2 * peaks.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
        peak_obs = N_SM * FP32_LANES * 2 * f_mhz * 1e6 / 1e12
        kernel_rate = world * samples_per_step * K / (dev_ms * 1e-3) / world  # per GPU
        achieved = kernel_rate * F / 1e12
        line['roofline'] = {
            'bound': 'fp32', 'achieved': achieved, 'peak': peak_max, 'unit': 'TFLOP/s', 'frac': achieved / peak_max,
            'traffic': ncu_traffic(wl) if (args.mode == 'fast' and args.pipeline == 'megakernel' and args.jit == 2) else None, 'kernel': 'pt_render_jit' if args.jit == 2 or sc.sdf_sources else 'pt_render_' + args.mode,
            'flops_per_sample_algorithmic': F, 'flops_counting_rule': 'SURVEY.md App. D v1 (source-level, as written in shader.comp)',
            'peak_basis': '148 SM x 128 FP32 lanes x 2 x %s sm_max_mhz (%s); FP32-issue roofline per SURVEY.md section 8d' % (
                peaks.get('sm_max_mhz', 1965.0), peak_kind),
            'frac_at_observed_clock': achieved / peak_obs,
            'hbm': {'achieved_gbs': W * H * 32 / (dev_ms / K * 1e-3) / 1e9, 'peak_gbs': peaks.get('hbm_gbs'),
                    'algorithmic_bytes_per_launch': W * H * 32},
            'oracle_counters_per_sample': {k: v / max(cnt['samples'], 1) for k, v in cnt.items()},
        }
        if not args.no_cpu_baseline:
            rate, threads, desc = cpu_rate(scene_name, W, H, pl, 12.0)
            line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port', 'sample': desc}
    except Exception as e:  # the baseline legs must never take the measurement down
        line['cpu_baseline_error'] = repr(e)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
