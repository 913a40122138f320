#!/usr/bin/env python
"""Headline benchmark: parallelgen IAF audio samples/s on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference

Workload (config.workload): BASELINE configs[2] — parallel WaveNet 4-flow IAF student
(parallel_wavenet.json, random init seed 12345), batch 8 x 7680 samples per GPU, synthetic
mel ~ U[0,1), noise drawn on the device.  One "step" = one full forward of that batch
(deconv stack, cond projections, 60 residual layers, 4 heads, quantise).  N > 1 shards
independent clips across ranks (weak scaling, 8 clips per GPU); NCCL is used once, to
broadcast the weight blob from rank 0, never in the timed loop.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'audio samples/sec (parallelgen IAF, 16 kHz)'
UNIT = 'samples/s'
CONFIGS = {'student': 'parallel_wavenet.json', 'clarinet': 'parallel_wavenet_gauss.json'}
# algorithmic bytes of the fused residual layer per (sample, layer):
# read l (64*4) + read cond (64*4) + write l (64*4)
LAYER_BYTES_PER_SAMPLE = 768
COND_FLOP_PER_SAMPLE_PLANE = 2 * 256 * 64


def load_hparams(name):
    from argparse import Namespace
    with open(os.path.join(ROOT, 'nsynth_wavenet_b200', 'config_jsons', CONFIGS[name])) as f:
        return Namespace(**json.load(f))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason sampling through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.ok or not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons)}


def cpu_reference_run(hp, batch, frames, steps, warmup, seed=12345):
    """Times the torch-CPU restatement of the reference's graph (oracle/torch_port.py) with
    every host thread.  Returns (samples_per_s, ms_per_step, cores, sample description)."""
    import torch
    from oracle import torch_port, wavenet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = O.init_student_weights(hp, seed=seed)
    port = torch_port.StudentPort(w, hp)
    rng = np.random.default_rng(54321)
    mel = rng.uniform(0, 1, (batch, frames, 80)).astype(np.float32)
    T = O.iaf_length(frames, hp)
    gauss = getattr(hp, 'loss_type', 'logistic') != 'logistic'
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if gauss:
            z = rng.standard_normal((batch, T)).astype(np.float32)
        else:
            z = O.logistic_from_uniform(rng.uniform(1e-5, 1 - 1e-5, (batch, T))).astype(np.float32)
        port.forward(mel, z, quantize=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    tot = float(np.sum(times))
    return batch * T * len(times) / tot, 1e3 * tot / len(times), cores, \
        '{} step(s) of {}x{} samples (full forward incl. noise draw), torch-CPU fp32, {} threads'.format(
            len(times), batch, T, cores)


def fastgen_bench(device, steps, hbm_peak):
    """Teacher WaveNet (wavenet_mol.json, 30 layers, random init) free-running generation at
    batch 1 through the persistent kernel; encoding resident in HBM, in-kernel RNG."""
    import torch
    from argparse import Namespace
    from nsynth_wavenet_b200 import FastgenEngine
    from oracle import wavenet_oracle as O
    with open(os.path.join(ROOT, 'nsynth_wavenet_b200', 'config_jsons', 'wavenet_mol.json')) as f:
        hp = Namespace(**json.load(f))
    w = O.init_teacher_weights(hp, seed=12345)
    eng = FastgenEngine(hp, w, device=device)
    g = torch.Generator(device='cpu').manual_seed(1)
    enc = (torch.rand((1, steps, 256), generator=g) * 2 - 1).to('cuda:%d' % device)
    eng.run_device(enc[:, :2048], seed=1)       # warm-up
    torch.cuda.synchronize()
    eng.run_device(enc, seed=2)
    torch.cuda.synchronize()
    ms = eng.last_timing()
    sps = steps / (ms * 1e-3)
    weight_bytes = 32 * 128 * 9224 * 4          # per-step streamed weight blocks (all CTAs)
    return {'metric': 'fastgen audio samples/sec (wavenet_mol.json, batch 1)', 'value': sps,
            'unit': UNIT, 'rtf': sps / 16000.0, 'steps': steps, 'ms': ms,
            'us_per_step': 1e3 * ms / steps,
            'weight_stream_gbs': weight_bytes * sps / 1e9,          # L2/HBM -> shared memory, all CTAs
            'weight_stream_frac_of_hbm': weight_bytes * sps / 1e9 / hbm_peak,
            # ncu capture of the same kernel (profiles/r01/fastgen_ncu_run56.json): 76.7 MB of DRAM reads per
            # sample, the rest of the 151 MB stream is served from the L2-resident (evict_last) blocks
            'dram_read_gbs_from_ncu_bytes': 76.7e6 * sps / 1e9,
            'note': 'includes the hoisted cond GEMM; weights stream L2/HBM -> smem every step, the blocks of '
                    'the first ~60 % of L2 worth of phases are loaded L2::evict_last (ncu: 49 % L2 hit rate)'}


def distill_bench(device, hbm_peak):
    """BASELINE configs[4] forward pieces on one GPU at the author's per-GPU batch (7 x 7680):
    student forward + teacher full-sequence forward on the student's output + 100-sample MoL
    cross-entropy (parallel_wavenet.py:361-402).  Secondary metric, reported not optimised."""
    import torch
    from argparse import Namespace
    from nsynth_wavenet_b200 import IAFEngine, TeacherEngine
    from oracle import wavenet_oracle as O
    cfgdir = os.path.join(ROOT, 'nsynth_wavenet_b200', 'config_jsons')
    with open(os.path.join(cfgdir, 'wavenet_mol.json')) as f:
        thp = Namespace(**json.load(f))
    shp = load_hparams('student')
    st = IAFEngine(shp, O.init_student_weights(shp, seed=12345), device=device)
    te = TeacherEngine(thp, O.init_teacher_weights(thp, seed=12345), device=device)
    dev = 'cuda:%d' % device
    mel = torch.rand((7, 39, 80), device=dev)
    res = None
    evs = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = st.forward_device(mel, None, seed=i, quantize=False)
        te_out = te.forward_device(out['x'], mel)
        res = te.mol_score(te_out, out['mean_tot'], out['scale_tot'], out['log_scale_tot'],
                           num_samples=100, seed=i)
        e1.record()
        torch.cuda.synchronize()
        evs.append(e0.elapsed_time(e1))
    ms = float(np.median(evs[1:]))
    return {'metric': 'distillation forward (student + teacher + 100-sample MoL CE), 7x7680 per GPU',
            'ms': ms, 'teacher_forward_ms': te.last_timing(), 'clips_per_s': 7 / (ms * 1e-3),
            'losses': res}


def run_reference(args, emit):
    """--impl reference: the reference's own CPU path restated (TF 1.x is not installable
    here), bounded sample, rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    hp = load_hparams(args.config)
    # bounded: 1 clip of 7680 samples per step keeps K=10,W=3 within a few minutes
    batch = args.ref_batch
    v, ms, cores, sample = cpu_reference_run(hp, batch, args.frames, args.steps, max(1, args.warmup))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'configs[2]: parallel WaveNet 4-flow IAF student ({}), '
                               '8x7680 per GPU; reference arm runs a bounded sample of it on host '
                               'cores'.format(CONFIGS[args.config])},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'rtf': v / 16000.0,
        'note': 'TensorFlow 1.x is absent: this is the fp32 torch-CPU restatement of the '
                'reference graph (oracle/torch_port.py), all host threads',
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='student', choices=list(CONFIGS))
    ap.add_argument('--batch', type=int, default=8, help='clips per GPU')
    ap.add_argument('--frames', type=int, default=39, help='mel frames per clip (39 -> 7680 samples)')
    ap.add_argument('--engine', default=None, choices=[None, 'ffma', 'tc', 'tc2', 'tc3'])
    ap.add_argument('--ref-batch', type=int, default=1)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fastgen', action='store_true')
    ap.add_argument('--no-distill', action='store_true')
    ap.add_argument('--fastgen-steps', type=int, default=32000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # the contract is ONE JSON line on stdout: anything libraries print (e.g. NCCL's version
    # banner) goes to stderr until the result line is emitted
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    if args.impl == 'reference':
        run_reference(args, emit)
        return

    import torch
    import torch.distributed as dist
    from nsynth_wavenet_b200 import IAFEngine, _lib
    from oracle import wavenet_oracle as O  # weights init + cpu_baseline leg only

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)

    hp = load_hparams(args.config)
    # ---- weights: rank 0 initialises, NCCL broadcasts one flat blob (the only collective) ----
    from nsynth_wavenet_b200 import parallel
    w0 = O.init_student_weights(hp, seed=12345)
    if rank != 0:  # only rank 0's values survive: proves the broadcast carries the weights
        w0 = {k: np.zeros_like(v) for k, v in w0.items()}
    weights = parallel.broadcast_weights(w0, device=dev)
    eng = IAFEngine(hp, weights, device=local_rank, engine=args.engine)

    B, F = args.batch, args.frames
    T = eng.length(F)
    rng = np.random.default_rng(54321 + rank)
    mel_h = torch.from_numpy(rng.uniform(0, 1, (B, F, 80)).astype(np.float32)).pin_memory()
    mel_d = mel_h.to(dev)
    out = {k: torch.empty((B, T), dtype=torch.float32, device=dev)
           for k in ('x', 'mean_tot', 'scale_tot', 'log_scale_tot')}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    lib = _lib.load()

    def step(i):
        eng.forward_device(mel_d, None, seed=1000 + i, quantize=True, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.nsw_kernel_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()                      # evict L2 between timed iterations (not timed)
        evs[i][0].record()
        step(i)
        evs[i][1].record()
    barrier()
    launches = lib.nsw_kernel_launch_count() - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    gpu_ms = sum(a.elapsed_time(b) for a, b in evs)
    total_ms = parallel.max_over_ranks(gpu_ms, device=dev)
    ms_per_step = total_ms / args.steps
    value = world * B * T * args.steps / (total_ms * 1e-3)

    # ---- e2e: the reference-facing host call, pinned host buffers, H2D + D2H inside ----
    x_h = np.empty((B, T), np.float32)
    x_pin = torch.from_numpy(x_h).pin_memory()
    mel_np = mel_h.numpy()
    for i in range(2):
        eng.forward_host(mel_np, None, seed=i, quantize=True)
    barrier()
    names5 = ('x', 'mean_tot', 'scale_tot', 'log_scale_tot', 'rand_input')
    e2e_s = 0.0
    for i in range(args.steps):
        flush.zero_()                      # same L2 hygiene as the device-timed loop (not timed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()           # the call is synchronous: returns with x in host memory
        _lib.check(lib.nsw_iaf_forward_host(eng._h, mel_h.data_ptr(), None, 2000 + i, B, F, 1,
                                            x_pin.data_ptr(), None, None, None, None))
        e2e_s += time.perf_counter() - t0
    e2e_value = world * B * T * args.steps / parallel.max_over_ranks(e2e_s, device=dev)

    # ---- roofline of the dominant kernel (iaf_layer_kernel), measured live with CUDA events
    #      on the launch stream inside the library (separate profiled passes) ----
    eng.set_profiling(True)
    prof = []
    for i in range(3):
        flush.zero_()
        step(i)
        torch.cuda.synchronize()
        prof.append(eng.last_timing())
    eng.set_profiling(False)
    stage = {k: float(np.median([p[k] for p in prof])) for k in prof[0]}
    n_layers = int(sum(hp.num_iaf_layers))
    n_planes = n_layers + len(hp.num_iaf_layers)
    layer_launch_ms = stage['layers'] / n_layers
    hbm_peak, tc_peak, peak_kind = measured_peaks()
    n_flows = len(hp.num_iaf_layers)
    fused = eng.engine == 'tc3'
    # algorithmic bytes of the 'layers' stage: 768 B per (sample, layer); engine tc3 also runs the
    # start conv (read x 4 B + write l 256 B) and the head (l 256 + cond 256 + ~40 B of x / totals) of every
    # flow inside the same kernel, so they are part of both the bytes and the time
    stage_bytes = B * T * (LAYER_BYTES_PER_SAMPLE * n_layers + (n_flows * (260 + 552) if fused else 0))
    n_launches = n_flows if fused else n_layers
    layer_gbs = stage_bytes / (stage['layers'] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (one 10-layer flow launch);
    # only valid for the shape and engine it was captured on
    traffic, traffic_src, traffic_alg = None, None, None
    ncu_json = os.path.join(ROOT, 'profiles', 'r01', 'tc3_ncu_run28.json')
    if fused and (B, T) == (8, 7680) and os.path.exists(ncu_json):
        with open(ncu_json) as f:
            caps = [c for c in json.load(f)['launches'] if 'iaf_flow_tc' in c['kernel']]
        if caps:
            cap = caps[0]
            traffic = (float(cap['dram__bytes_read.sum']['value']) + float(cap['dram__bytes_write.sum']['value'])) * 1e6
            traffic_alg = B * T * (LAYER_BYTES_PER_SAMPLE * 10 + 812)
            traffic_src = ('profiles/r01/tc3_ncu_run28.json: dram__bytes_read+write of one iaf_flow_tc_kernel launch '
                           '(10-layer flow incl. start conv and head); traffic_launch_algorithmic_bytes is the same '
                           'launch under the 768 B model')
    cond_tflops = COND_FLOP_PER_SAMPLE_PLANE * n_planes * B * T / (stage['cond'] * 1e-3) / 1e12

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {
            'workload': ('configs[2]: parallel WaveNet 4-flow IAF student ({}), batch {}x{} samples '
                         'per GPU, synthetic mel, device-drawn noise' if args.config == 'student' else
                         'configs[3] per-GPU share: ClariNet Gaussian IAF ({}), batch {}x{} samples per GPU, '
                         'synthetic mel, device-drawn noise').format(CONFIGS[args.config], B, T),
            'clips_per_gpu': B, 'samples_per_clip': T, 'engine': eng.engine,
            'l2': 'flushed between timed iterations (256 MB write)', 'parallelism': 'clips x{}'.format(world),
        },
        'rtf': value / 16000.0,
        'rtf_per_gpu': value / 16000.0 / world,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(B * F * 80 * 4),
                'd2h_bytes_per_step': int(B * T * 4)},
        'gpu_launches': int(launches),
        'clocks': sampler.summary(),
        'roofline': {
            'kernel': {'tc3': 'iaf_flow_tc_kernel', 'tc2': 'iaf_layer_tc_kernel'}.get(eng.engine, 'iaf_layer_kernel'), 'bound': 'hbm', 'achieved': layer_gbs, 'peak': hbm_peak,
            'unit': 'GB/s', 'frac': layer_gbs / hbm_peak, 'traffic': traffic,
            'traffic_launch_algorithmic_bytes': traffic_alg, 'traffic_source': traffic_src,
            'unit_of_launch': ('one flow (start conv + all residual layers + head) over the whole batch; '
                               'bytes_per_launch and launch_ms are the averages over the flows'
                               if fused else 'one residual layer over the whole batch'),
            'peak_source': peak_kind + ' (MEASURED_PEAKS.json hbm_gbs)',
            'bytes_per_launch': stage_bytes / n_launches, 'launch_ms': stage['layers'] / n_launches,
            'share_of_step': stage['layers'] / stage['total'],
        },
        'roofline_cond_gemm': {
            'kernel': 'cond_proj_tc_kernel' if eng.engine == 'tc3' else 'conv_gemm (cond projections)', 'bound': 'tensor', 'achieved': cond_tflops,
            'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': cond_tflops / tc_peak,
            'note': 'algorithmic fp32-equivalent flops; the tcgen05 engine issues 3 fp16 MMAs per product',
        },
        'stage_ms': stage,
    }

    # ---- secondary metric: autoregressive fastgen (BASELINE configs[1]), batch 1, rank 0 ----
    if rank == 0 and not args.no_fastgen:
        try:
            line['fastgen'] = fastgen_bench(local_rank, args.fastgen_steps, hbm_peak)
        except Exception as ex:  # reported, never silently dropped
            line['fastgen'] = {'error': str(ex)[:300]}

    if rank == 0 and not args.no_distill:
        try:
            line['distill'] = distill_bench(local_rank, hbm_peak)
        except Exception as ex:
            line['distill'] = {'error': str(ex)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample (about cpu-seconds of work): single-clip forwards
        v1, ms1, cores, _ = cpu_reference_run(hp, 1, F, 1, 1)
        reps = int(max(1, min(20, args.cpu_seconds / (ms1 * 1e-3))))
        v, ms, cores, sample = cpu_reference_run(hp, 1, F, reps, 0)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': sample}
    elif rank == 0:
        line['cpu_baseline'] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
