#!/usr/bin/env python3
"""bench.py -- spectral path samples/s of libpt_cuda on BASELINE.json's configs (driver contract: see DESIGN.md).

  python bench.py --gpus 1 --steps K --warmup W            our arm, one B200
  torchrun ... bench.py --gpus N --steps K --warmup W      our arm, one rank per GPU (sample-split, one NCCL reduce)
  python bench.py --impl reference ...                     the reference's own shader.comp on the host cores (oracle/_ref)

A "step" is one dispatch of the hot path: --spf samples per pixel over the whole frame of the workload.  The headline
workload (`value`, `e2e`, `roofline`, `cpu_baseline`) is BASELINE config 5 -- scenes/scene10.json (menger SDF) at
3840x2160, pathLength 5, camera shot 1: the one config BASELINE.json ties to 1/2/4/8 GPUs, and it fits one GPU -- and
the other BASELINE configs are measured in the same run with fewer steps and reported under `config.per_workload`.
1 sample = one Scene() call (shader.comp:1446-1490): one 4-wavelength hero bundle through the camera lens and the
whole path.  `value` is device time (CUDA events on the library's stream, inputs resident in HBM, L2 flushed between
steps); `e2e` is the same metric through the C ABI with host buffers: at N = 1 scene block upload + dispatch + read-back
of the XYZ image into pinned host memory every step; at N > 1 every rank uploads and dispatches its slice of the
samples, then ONE reduce, finalize and ONE read-back on rank 0.  Data is synthetic in the sense of the contract: the
reference's own shipped scene files, no external assets.
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
HEADLINE = 'cfg5_scene10_4k'
BASELINE_WORKLOADS = ['cfg1_scene0_512', 'cfg2_scene1_1080p', 'cfg3_scene9_mandelbulb_1080p', 'cfg4a_scene10_menger_1080p_pl32',
                      'cfg4b_scene8_terrain_1080p_pl32', 'cfg5_scene10_4k']


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


_T_PROCESS = time.perf_counter()


def emit(line):
    line.setdefault('wall_s_process', round(time.perf_counter() - _T_PROCESS, 2))  # everything: imports, JIT, CPU legs, parity
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
    `ncu --set full` capture of the same bench command at 64 samples per step (profiles/<round>/ncu_summary_<cfg>.json):
    the texel read-modify-write plus, with option pregen, the 32-byte records the render kernel reads.  Returns (bytes or
    None, the file it came from)."""
    for sub in ('r02_final2', 'r02_final', 'r02_gpu1', 'r01_final4', 'r01_final'):
        path = os.path.join(ROOT, 'profiles', sub, 'ncu_summary_%s.json' % wl.split('_')[0])
        try:
            with open(path) as f:
                d = json.load(f)
            if d.get('workload') == wl:
                return d['traffic_bytes_per_launch'], os.path.relpath(path, ROOT)
        except (OSError, ValueError, KeyError):
            pass
    return None, None


def host_threads():
    """The host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which is about their
    torch thread pools and not about a CPU baseline: it is ignored here (round 1's SCALE records timed the reference arm
    on one core for that reason)."""
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return max(os.cpu_count() or 1, 1)


def oracle_for(scene_name, count=False, threads=0):
    from oracle import oracle, pack
    scene = pack.load_scene(scene_file(scene_name))
    return oracle.Oracle(pack.pack_ubo(scene), pack.sdf_sources(scene), count=count, threads=threads), scene


class CpuArm:
    """The path on the host cores: oracle/_ref -- the reference's own src/shader.comp and packer compiled for the CPU
    (kind "reference") -- when it was built (in the container that holds /root/reference; the objects travel with the
    repo), else the oracle port (kind "port").  One dispatch() = one vkCmdDispatch over a strided subset of the rows."""

    def __init__(self, scene_name, prefer_ref=True):
        from oracle import pack
        self.threads = host_threads()
        self.kind = 'port'
        self.scene_name = scene_name
        self._ref = None
        if prefer_ref:
            try:
                from oracle import ref
                if ref.available():
                    self._ref = ref.RefScene(scene_file(scene_name))
                    self._ref.shader()
                    self.kind = 'reference'
            except Exception as e:  # an unusable _ref must not take the bench down: say so and time the port
                self._ref = None
                self.note = 'oracle/_ref unusable: %r' % (e,)
        if self._ref is None:
            self._o, self._scene = oracle_for(scene_name, threads=self.threads)
            self._pack = pack

    def dispatch(self, width, height, spf, path_length, dispatch_index, row_step):
        img = np.zeros((height, width, 4), dtype=np.float32)
        t0 = time.perf_counter()
        if self._ref is not None:
            push = self._ref.push(width, height, dispatch_index * spf, dispatch_index * spf, spf, path_length)
            self._ref.dispatch(push, img, 0, row_step, None, self.threads)
        else:
            p = self._pack.pack_params(self._scene, 1, width, height, spf, path_length, dispatch=dispatch_index)
            self._o.dispatch(p, img, 0, row_step)
        return time.perf_counter() - t0, len(range(0, height, row_step))

    def rate(self, width, height, path_length, target_s, dispatch_index=1):
        """A bounded sample of the workload sized for about target_s seconds: a strided subset of the rows of the
        full-resolution frame (or, when one whole frame is too quick, several samples per pixel of the whole frame).
        Returns (samples/s, description)."""
        step = max(height // max(2 * self.threads, 8), 1)      # probe: a few rows per thread
        dt, rows = self.dispatch(width, height, 1, path_length, dispatch_index, step)
        rate = rows * width / max(dt, 1e-6)
        want = rate * target_s
        spf = int(min(max(want // (width * height), 1), 64))
        rows = int(min(max(want // (width * spf), 2 * self.threads), height))
        step = max(height // rows, 1)
        dt, rows = self.dispatch(width, height, spf, path_length, dispatch_index, step)
        return rows * width * spf / dt, '%d of %d rows (every %d-th) of %s at %dx%d, %d spp, pathLength %d: %.2f s on %d threads' % (
            rows, height, step, self.scene_name, width, height, spf, path_length, dt, self.threads)


def counters_for(scene_name, width, height, path_length, rows=24):
    from oracle import pack
    o, scene = oracle_for(scene_name, count=True, threads=host_threads())
    p = pack.pack_params(scene, 1, width, height, 1, path_length)
    img = np.zeros((height, width, 4), dtype=np.float32)
    cnt = o.dispatch(p, img, 0, max(height // rows, 1))
    counts = tuple(int(v) for v in o.ubo[:6])
    return cnt, counts


def run_reference(args, wl):
    """--impl reference: the reference's own implementation of the path on the host cores, all of them -- oracle/_ref
    (src/shader.comp compiled over the vendored glm, src/pathtracer.cpp's loader and packer), else the oracle port."""
    scene_name, W, H, spp_cfg, pl, c_sdf, c_mat = WORKLOADS[wl]
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    arm = CpuArm(scene_name)
    total = args.steps + args.warmup
    per_step = min(max(90.0 / max(total, 1), 0.05), 5.0)
    rates, desc = [], ''
    t_all = time.perf_counter()
    for i in range(total):
        rate, desc = arm.rate(W, H, pl, per_step, dispatch_index=1 + i)
        if i >= args.warmup:
            rates.append(rate)
    wall = time.perf_counter() - t_all
    value = float(len(rates) / sum(1.0 / r for r in rates)) if rates else 0.0
    what = ('the reference\'s src/shader.comp + packer compiled for the CPU (oracle/_ref)' if arm.kind == 'reference'
            else 'CPU oracle port of shader.comp (oracle/_ref was not built)')
    line = {
        'impl': 'reference', 'metric': 'spectral path samples/sec', 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(total, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic (reference scene files shipped in scenes/)',
        'config': {'workload': wl, 'scene': os.path.relpath(scene_file(scene_name), ROOT), 'width': W, 'height': H, 'path_length': pl,
                   'spp_of_config': spp_cfg, 'note': what + '; each step = a strided-row sample of the frame'},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': arm.threads, 'kind': arm.kind, 'sample': desc},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if arm.threads == 1 and (os.cpu_count() or 1) > 1:
        line['warning'] = 'reference arm ran on ONE core of a multi-core host (affinity mask): ratios against it are not comparable'
    emit(line)
    return 0


class Bench:
    """One process = one GPU of the job.  measure() times one workload: K steps device-timed, then the same K steps
    end to end through host buffers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import pathtracer_b200 as pt
        self.torch, self.dist, self.pt, self.args = torch, dist, pt, args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; libpt_cuda has no CPU fallback (use --impl reference for the CPU arm)')
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
        self.flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        self.options = {}
        for kv in args.opt or []:
            k, v = kv.split('=', 1)
            self.options[k] = int(v)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor([float(v) for v in vals], device='cuda', dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def renderer(self, mode=None):
        pt, a = self.pt, self.args
        m = (pt.MODE_FAST if a.mode == 'fast' else pt.MODE_STRICT) if mode is None else mode
        r = pt.Renderer(device=self.local_rank, mode=m, jit=a.jit,
                        pipeline=pt.PIPE_WAVEFRONT if a.pipeline == 'wavefront' else pt.PIPE_MEGAKERNEL, options=self.options)
        if a.bvh_min is not None:
            r.set_bvh(a.bvh_min)
        return r

    def measure(self, wl, K, Wm, sampler=None):
        torch, dist, pt = self.torch, self.dist, self.pt
        world, rank = self.world, self.rank
        scene_name, W, H, spp_cfg, pl, c_sdf, c_mat = WORKLOADS[wl]
        spf = self.args.spf
        sc = pt.Scene.load(scene_file(scene_name))
        ubo = sc.pack_ubo()
        params = sc.pack_params(1, W, H, spf, pl)
        r = self.renderer()
        t0 = time.perf_counter()
        r.set_scene(ubo, sc.sdf_sources)
        compile_s = time.perf_counter() - t0
        image = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
        r.bind_image(image)
        ext = torch.cuda.ExternalStream(r.stream)
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
            if self.flush is not None:
                with torch.cuda.stream(ext):
                    self.flush.fill_(rank & 0xFF)

        def reduce_and_finalize(total_spp):
            """the path's one exchange step: NCCL reduce of the per-GPU sum images to rank 0, finalize there.  The
            library's stream is not torch's: order the two explicitly (r.sync before, synchronize after)."""
            r.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.reduce(image, dst=0, op=dist.ReduceOp.SUM)
            e1.record()
            torch.cuda.synchronize()
            if rank == 0:
                r.finalize(params, total_spp)
                r.sync()
            return e0.elapsed_time(e1)

        for i in range(Wm):
            step(i)
            flush_l2()
        r.sync()
        r.kernel_time()
        self.barrier()
        if sampler is not None:
            sampler.mark_begin()
        wall0 = time.perf_counter()
        dev_ms, launches = 0.0, 0
        for i in range(Wm, Wm + K):
            step(i)
            ms, n = r.kernel_time()  # CUDA events on the library's stream; synchronises
            dev_ms += ms
            launches += n
            flush_l2()
        reduce_ms = reduce_and_finalize(world * (K + Wm) * spf) if sum_mode else 0.0
        self.barrier()
        wall = time.perf_counter() - wall0
        if sampler is not None:
            sampler.mark_end()
        total_ms, dev_ms_max = self.max_over_ranks(dev_ms + reduce_ms, dev_ms)
        samples_per_step = W * H * spf
        value = world * samples_per_step * K / (total_ms * 1e-3)

        # ---- e2e: the call a user makes, host buffers in, host buffers out ------------------------------------------
        hosts = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2 if not sum_mode else 1)]
        hosts_np = [h.numpy() for h in hosts]
        r.clear()
        self.barrier()
        r.read_xyz_async(hosts_np[0])               # untimed: creates the copy stream and the snapshot buffer
        r.read_wait()
        self.barrier()
        t0 = time.perf_counter()
        if not sum_mode:
            for i in range(K):
                r.set_scene(ubo, sc.sdf_sources)       # h2d: the 16 388-byte uniform block (kernel cache hit)
                step(i)                                # the 88-byte push block travels with the launch
                r.read_xyz_async(hosts_np[i & 1])      # d2h: W*H*16 bytes into pinned host memory, overlapping the next step
            r.read_wait()                              # every step's image has landed on the host
            r.sync()
            d2h_per_step = W * H * 16
        else:
            for i in range(K):
                r.set_scene(ubo, sc.sdf_sources)
                step(i)
            reduce_and_finalize(world * K * spf)       # one exchange, then the image goes to the host once, on rank 0
            if rank == 0:
                r.read_xyz_async(hosts_np[0])
                r.read_wait()
            d2h_per_step = W * H * 16 / K              # one read-back of the final image, spread over the K steps
        e2e_s = time.perf_counter() - t0
        (e2e_s,) = self.max_over_ranks(e2e_s)
        e2e_value = world * samples_per_step * K / e2e_s
        r.kernel_time()

        out = {'workload': wl, 'scene_name': scene_name, 'W': W, 'H': H, 'pl': pl, 'spp_cfg': spp_cfg, 'c_sdf': c_sdf, 'c_mat': c_mat,
               'value': value, 'ms_per_step': total_ms / K, 'dev_ms': dev_ms_max, 'reduce_ms': reduce_ms, 'launches': int(launches),
               'e2e_value': e2e_value, 'h2d': 16388 + 88, 'd2h': d2h_per_step, 'compile_s': compile_s, 'bvh_active': r.bvh_active,
               'wall': wall, 'K': K, 'Wm': Wm, 'sched': r.get_option('sched'), 'has_sdf': bool(sc.sdf_sources)}

        # ---- N > 1: the reduced image against a one-rank render of the same sample range -----------------------------
        if sum_mode and self.args.multi_parity:
            r.clear()
            r.sync()
            r.dispatch_sum(params, rank * spf, spf)
            reduce_and_finalize(world * spf)
            split = image.clone() if rank == 0 else None
            self.barrier()
            if rank == 0:
                r.clear()
                for g in range(world):
                    r.dispatch_sum(params, g * spf, spf)
                r.finalize(params, world * spf)
                r.sync()
                torch.cuda.synchronize()
                a, b = split[..., :3].double(), image[..., :3].double()
                scale = float(b.abs().max())
                err = (a - b).abs()
                tol = 1e-5 * b.abs() + 1e-6 * scale
                out['multi_parity'] = {'ok': bool((err <= tol).all()), 'max_abs_err_over_scale': float(err.max()) / max(scale, 1e-30),
                                       'fraction_within_rtol_1e-5': float((err <= tol).double().mean()), 'samples_per_rank': spf,
                                       'what': 'sample-split over %d ranks + reduce + finalize vs rank 0 rendering the same %d sample indices alone' % (world, world * spf)}
            self.barrier()
        r.close()
        del image
        return out

    def time_to_rel_rmse(self, wl, value):
        """Second half of BASELINE.json's metric, on rank 0 at N = 1: relRMSE of FAST renders against a STRICT reference
        with disjoint sample indices, at the workload's resolution / 8 (the statistic does not depend on the pixel count;
        the reference at full size would take minutes in strict mode).  relMSE(spp) is fitted as a * (1/spp + 1/ref_spp);
        the time to a target is spp_needed x pixels of the config / measured samples per second."""
        torch, pt = self.torch, self.pt
        scene_name, W, H, spp_cfg, pl, _, _ = WORKLOADS[wl]
        w, h = max(W // 8, 64), max(H // 8, 48)
        ref_spp = 16384
        sc = pt.Scene.load(scene_file(scene_name))
        ubo = sc.pack_ubo()
        p = sc.pack_params(1, w, h, 64, pl)
        img = torch.zeros((h, w, 4), dtype=torch.float32, device='cuda')

        def render(mode, first, spp):
            r = self.renderer(mode)
            r.set_scene(ubo, sc.sdf_sources)
            r.bind_image(img)
            img.zero_()
            torch.cuda.synchronize()
            r.kernel_time()
            s = first
            while s < first + spp:
                n = min(256, first + spp - s)
                r.dispatch_sum(p, s, n)
                s += n
            r.finalize(p, spp)
            ms, _ = r.kernel_time()
            out = img.clone()
            r.close()
            return out, ms * 1e-3

        ref, ref_s = render(pt.MODE_STRICT, 1 << 20, ref_spp)
        rr = ref[..., :3].double()
        eps = (0.01 * rr.mean()) ** 2
        pts = []
        for spp in (16, 64, 256, 1024):
            im, _ = render(pt.MODE_FAST if self.args.mode == 'fast' else pt.MODE_STRICT, 0, spp)
            mse = float(torch.mean((im[..., :3].double() - rr) ** 2 / (rr ** 2 + eps)))
            pts.append((spp, mse))
        a_i = [mse / (1.0 / spp + 1.0 / ref_spp) for spp, mse in pts]
        a = float(np.median(a_i))
        targets = {}
        for th in (0.2, 0.1, 0.05):
            need = a / (th * th)
            targets[str(th)] = {'spp': need, 'seconds_at_config_resolution': need * W * H / value}
        return {'reference': 'strict mode, %d spp, sample indices from 2^20 (disjoint), %dx%d (config resolution / 8), %.2f s' % (ref_spp, w, h, ref_s),
                'rel_rmse_points': [{'spp': s, 'rel_rmse': float(np.sqrt(m))} for s, m in pts],
                'rel_mse_times_spp': a, 'fit_spread': float(max(a_i) / min(a_i)), 'targets': targets,
                'note': 'relMSE(spp) = a (1/spp + 1/ref_spp); a spread near 1 means no bias floor between fast and strict'}


def roofline_for(m, peaks, peak_kind, fp32_peak, clocks, args):
    """FP32-issue roofline of one measured workload: algorithmic flops (oracle call counters x SURVEY App. D constants)
    over the device time of its kernel launches."""
    wl = m['workload']
    if wl.startswith('bvh_'):
        # the source-level count is the reference's brute-force scan; the tree skips most of it, so dividing that count
        # by the tree kernel's time is not a fraction of anything (round 1 printed 0.96 here): no roofline for these
        return {'bound': 'fp32', 'achieved': None, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': None, 'traffic': None,
                'note': 'no algorithmic numerator for the BVH workloads: the counting rule describes the scan the tree avoids'}
    cnt, counts = counters_for(m['scene_name'], m['W'], m['H'], m['pl'])
    F = flops_per_sample(cnt, counts, m['c_sdf'], m['c_mat'])
    derived_max = N_SM * FP32_LANES * 2 * peaks.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
    peak = fp32_peak or derived_max
    per_gpu_rate = m['W'] * m['H'] * args.spf * m['K'] / (m['dev_ms'] * 1e-3)
    achieved = per_gpu_rate * F / 1e12
    traffic, traffic_src = (ncu_traffic(wl) if (args.mode == 'fast' and args.pipeline == 'megakernel' and args.jit == 2) else (None, None))
    return {
        'bound': 'fp32', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
        'traffic': traffic, 'traffic_source': traffic_src,
        'kernel': 'pt_render_jit' if args.jit == 2 or m['has_sdf'] else 'pt_render_' + args.mode,
        # option "pregen" (auto for scenes with a cyclide): a step is pt_gen_jit (camera rays, 32-byte records) + pt_render_jit;
        # `achieved` divides by the device time of BOTH, the records add 64 B per sample (written once, read once) to the HBM side
        'kernels_per_step': m['launches'] / max(m['K'], 1),
        'flops_per_sample_algorithmic': F, 'flops_counting_rule': 'SURVEY.md App. D v1 (source-level, as written in shader.comp)',
        'peak_basis': ('measured in this run: pt_fp32_peak, 16 independent FFMA chains per thread on every SM, CUDA events' if fp32_peak else
                       'derived: 148 SM x 128 FP32 lanes x 2 x %s sm_max_mhz (%s)' % (peaks.get('sm_max_mhz', 1965.0), peak_kind)),
        'peak_derived_at_max_clock': derived_max, 'frac_of_derived_peak': achieved / derived_max,
        'hbm': (lambda rec: {'achieved_gbs': (m['W'] * m['H'] * 32 + rec) / (m['dev_ms'] / m['K'] * 1e-3) / 1e9, 'peak_gbs': peaks.get('hbm_gbs'),
                             'algorithmic_bytes_per_launch': m['W'] * m['H'] * 32 + rec,
                             'of_which_pregen_records': rec})((m['W'] * m['H'] * args.spf * 64) if m['launches'] > m['K'] else 0),
        'oracle_counters_per_sample': {k: v / max(cnt['samples'], 1) for k, v in cnt.items()},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument('--per-workload', dest='per_workload', action='store_true', default=None,
                    help='also measure the other BASELINE configs (default: yes when --workload is the headline)')
    ap.add_argument('--no-per-workload', dest='per_workload', action='store_false')
    ap.add_argument('--spf', type=int, default=64, help='samples per pixel per step (one dispatch; the sample pool holds 32 x spf items per warp)')
    ap.add_argument('--mode', default='fast', choices=['fast', 'strict'])
    ap.add_argument('--jit', type=int, default=2, help='0 static kernels, 1 NVRTC for SDF scenes only, 2 NVRTC scene-specialised')
    ap.add_argument('--pipeline', default='megakernel', choices=['megakernel', 'wavefront'])
    ap.add_argument('--bvh-min', type=int, default=None, help='bounded primitives from which the BVH replaces the scan (0: never)')
    ap.add_argument('--opt', action='append', help='tuning option key=value (pt_set_option), e.g. --opt sched=8; repeatable')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-rmse', action='store_true')
    ap.add_argument('--no-multi-parity', dest='multi_parity', action='store_false', default=True)
    ap.add_argument('--no-flush', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least three untimed steps
    quiet_stdout()
    wl = args.workload
    if args.impl == 'reference':
        return run_reference(args, wl)
    per_workload = args.per_workload if args.per_workload is not None else (wl == HEADLINE)

    b = Bench(args)
    torch, dist, world, rank = b.torch, b.dist, b.world, b.rank
    K, Wm, spf = args.steps, args.warmup, args.spf
    try:
        dev_uuid = str(torch.cuda.get_device_properties(b.local_rank).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(b.local_rank, dev_uuid)
    if rank == 0:
        sampler.start()
    m = b.measure(wl, K, Wm, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None

    others = []
    if per_workload:
        for w2 in BASELINE_WORKLOADS:
            if w2 != wl:
                others.append(b.measure(w2, min(K, 6), 3))

    fp32_peak = None
    rmse = None
    if rank == 0:
        try:
            r = b.renderer()
            fp32_peak, _ = r.fp32_peak(5)
            r.close()
        except Exception:
            fp32_peak = None
        if world == 1 and not args.no_rmse and args.pipeline == 'megakernel':
            try:
                rmse = b.time_to_rel_rmse(wl, m['value'])
            except Exception as e:
                rmse = {'error': repr(e)}
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = load_peaks()
    W, H = m['W'], m['H']
    line = {
        'metric': 'spectral path samples/sec', 'value': m['value'], 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': m['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic (reference scene files shipped in scenes/, no external assets)',
        'config': {'workload': wl, 'scene': os.path.relpath(scene_file(m['scene_name']), ROOT), 'width': W, 'height': H, 'spf_per_step': spf,
                   'spp_timed': K * spf * world, 'spp_of_config': m['spp_cfg'], 'path_length': m['pl'], 'shot': 1, 'mode': args.mode,
                   'jit': args.jit, 'pipeline': args.pipeline, 'driver_sched': m['sched'], 'options': b.options,
                   'closest_hit': 'bvh' if m['bvh_active'] else 'scan',
                   'l2': 'not flushed' if args.no_flush else 'flushed between steps (256 MiB fill)',
                   'parallelism': 'sample-split x%d + 1 NCCL reduce' % world if world > 1 else 'single GPU',
                   'kernel_compile_s': round(m['compile_s'], 3)},
        'e2e': {'value': m['e2e_value'], 'unit': 'samples/s', 'h2d_bytes_per_step': m['h2d'], 'd2h_bytes_per_step': m['d2h'], 'steps': K,
                'what': ('every step: scene block upload, dispatch, read-back of the XYZ image into pinned host memory' if world == 1 else
                         'every step: scene block upload + dispatch of the rank\'s sample slice; then one NCCL reduce, finalize and one read-back on rank 0')},
        'gpu_launches': m['launches'],
        'clocks': clocks,
        'wall_s_timed_region': m['wall'],
        'reduce_ms': m['reduce_ms'],
    }
    if 'multi_parity' in m:
        line['multi_parity'] = m['multi_parity']
    if rmse is not None:
        line['time_to_rel_rmse'] = rmse
    # ---- roofline + CPU baseline (rank 0; oracle/ is the checker and the baseline, never the thing measured above) -----
    try:
        from oracle import oracle as _o
        _o.build()
        line['roofline'] = roofline_for(m, peaks, peak_kind, fp32_peak, clocks, args)
        if not args.no_cpu_baseline and world == 1:
            arm = CpuArm(m['scene_name'])
            rate, desc = arm.rate(W, H, m['pl'], 12.0)
            line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': arm.threads, 'kind': arm.kind, 'sample': desc}
        if others:
            pw = {}
            for o in others:
                rf = roofline_for(o, peaks, peak_kind, fp32_peak, clocks, args)
                e = {'value': o['value'], 'ms_per_step': o['ms_per_step'], 'steps': o['K'], 'e2e': o['e2e_value'], 'gpu_launches': o['launches'],
                     'width': o['W'], 'height': o['H'], 'path_length': o['pl'], 'driver_sched': o['sched'],
                     'roofline_frac': rf['frac'], 'tflops_algorithmic': rf['achieved'], 'flops_per_sample_algorithmic': rf.get('flops_per_sample_algorithmic')}
                if 'multi_parity' in o:
                    e['multi_parity_ok'] = o['multi_parity']['ok']
                if not args.no_cpu_baseline and world == 1:
                    arm = CpuArm(o['scene_name'])
                    rate, desc = arm.rate(o['W'], o['H'], o['pl'], 2.0)
                    e['cpu_baseline'] = {'value': rate, 'cores': arm.threads, 'kind': arm.kind, 'sample': desc}
                pw[o['workload']] = e
            line['config']['per_workload'] = pw
    except Exception as e:  # the baseline legs must never take the measurement down
        line['cpu_baseline_error'] = repr(e)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
