#!/usr/bin/env python
"""Headline benchmark: parallelgen IAF audio samples/s on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference, same config

Workload (config.workload): BASELINE configs[2] -- parallel WaveNet 4-flow IAF student
(parallel_wavenet.json, random init seed 12345), batch 8 x 7680 samples per GPU, synthetic
mel ~ U[0,1), noise drawn on the device.  One "step" = one full forward of that batch
(deconv stack, cond projections, 60 residual layers, 4 heads, quantise).  N > 1 shards
independent clips across ranks (weak scaling, 8 clips per GPU); NCCL is used once, to
broadcast the weight blob from rank 0, never in the timed loop.

Beside the contract keys the line carries: `sustained` (the same step looped for >= 2 s), `e2e` (C-ABI host call,
pinned buffers), `e2e_python` (wavenet.parallelgen.synthesis: pageable arrays, checkpoint on disk, wav files written),
`roofline` (+ `frac_model_hbm`, `frac_dram`, `frac_tensor`), `roofline_cond_gemm`, `roofline_teacher`, `fastgen`
(configs[1]: latency engine at batch 1, batched engine at batch 8, its own e2e and cpu_baseline), `clarinet`
(configs[3] per-GPU share) and `distill` (configs[4] forward), the last two timed as max over ranks at every N.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
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
# algorithmic flops (2 * MAC) per (sample, layer): 3-tap 64->64 conv + 32->64 residual 1x1
LAYER_FLOP_PER_SAMPLE = 2 * (3 * 64 * 64 + 32 * 64)
HEAD_FLOP_PER_SAMPLE = 2 * (64 * 64 + 2 * 64)        # out1 + out2_mean / out2_scale, per (sample, flow)
COND_FLOP_PER_SAMPLE_PLANE = 2 * 256 * 64
SPLIT_PRODUCTS = 3                                   # hi*hi + hi*lo + lo*hi fp16 MMAs per fp32-grade product
FASTGEN_ALGORITHMIC_BYTES = 29697024 * 4             # weights touched per autoregressive step
TEACHER_FLOP_PER_SAMPLE = 67.4e6                     # full-sequence teacher, per output sample


def load_hparams(name):
    from argparse import Namespace
    fname = CONFIGS.get(name, name)
    with open(os.path.join(ROOT, 'nsynth_wavenet_b200', 'config_jsons', fname)) as f:
        return Namespace(**json.load(f))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


def sustained_tc_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as f:
            return float(json.load(f).get('bf16_tflops_sustained') or 0.0) or None
    except (OSError, ValueError):
        return None


def ncu_capture(name):
    """Committed ncu --set full summary of a kernel (profiles/r02 first, else r01): dict of metrics or None."""
    for rnd, fname in (('r02', 'flow_cond_ncu.json'), ('r01', 'tc3_ncu_run28.json')):
        p = os.path.join(ROOT, 'profiles', rnd, fname)
        if not os.path.exists(p):
            continue
        with open(p) as f:
            caps = [c for c in json.load(f)['launches'] if name in c['kernel']]
        if caps:
            return caps[0], 'profiles/{}/{}'.format(rnd, fname)
    return None, None


def metric_value(cap, key):
    v = cap.get(key)
    if isinstance(v, dict):
        v = v.get('value')
    return None if v is None else float(v)


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason sampling through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.power = []
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
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
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

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.summary()

    def summary(self):
        if not self.ok or not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        out = {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
               'reasons': sorted(self.reasons), 'samples': len(self.samples)}
        if self.power:
            out['power_w_max'] = float(np.max(self.power))
        return out


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/): cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(hp, batch, frames, steps, warmup, seed=12345):
    """Times the torch-CPU restatement of the reference's graph (oracle/torch_port.py) with
    every host thread.  Returns (samples_per_s, ms_per_step, cores, sample description)."""
    import torch
    from oracle import torch_port
    from nsynth_wavenet_b200.weights_init import init_student_weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = init_student_weights(hp, seed=seed)
    port = torch_port.StudentPort(w, hp)
    rng = np.random.default_rng(54321)
    mel = rng.uniform(0, 1, (batch, frames, 80)).astype(np.float32)
    T = (frames * 200 // 512) * 512
    gauss = getattr(hp, 'loss_type', 'logistic') != 'logistic'
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if gauss:
            z = rng.standard_normal((batch, T)).astype(np.float32)
        else:
            u = rng.uniform(1e-5, 1 - 1e-5, (batch, T))
            z = (np.log(u) - np.log1p(-u)).astype(np.float32)
        port.forward(mel, z, quantize=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    tot = float(np.sum(times))
    return batch * T * len(times) / tot, 1e3 * tot / len(times), cores, \
        '{} step(s) of {}x{} samples (full forward incl. noise draw), torch-CPU fp32, {} threads'.format(
            len(times), batch, T, cores)


def cpu_fastgen_run(steps):
    """the reference's per-sample loop (fastgen.py:156-168) restated on torch-CPU,
    wavenet_mol.json, batch 1, free-running MoL sampling."""
    import torch
    from oracle import torch_port
    from nsynth_wavenet_b200.weights_init import init_teacher_weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = load_hparams('wavenet_mol.json')
    port = torch_port.FastgenPort(init_teacher_weights(hp, seed=12345), hp, 1)
    rng = np.random.default_rng(1)
    enc = rng.uniform(-1, 1, (1, steps + 64, 256)).astype(np.float32)
    port.run(enc, 64)
    port.reset()
    t0 = time.perf_counter()
    port.run(enc, steps)
    dt = time.perf_counter() - t0
    return {'value': steps / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'rtf': steps / dt / 16000.0,
            'sample': '{} autoregressive steps, batch 1, one Python iteration per sample like fastgen.py:156, '
                      'torch-CPU fp32, {} threads'.format(steps, cores)}


def run_reference(args, emit):
    """--impl reference: the reference's own CPU path restated (TF 1.x is not installable
    here), same config as the b200 arm, rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    hp = load_hparams(args.config)
    batch = args.ref_batch
    v, ms, cores, sample = cpu_reference_run(hp, batch, args.frames, args.steps, max(1, args.warmup))
    T = (args.frames * 200 // 512) * 512
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.config, batch, T, None, args.gpus),
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'rtf': v / 16000.0,
        'note': 'TensorFlow 1.x is absent: this is the fp32 torch-CPU restatement of the '
                'reference graph (oracle/torch_port.py), all host threads, one process (rank 0)',
    }
    emit(line)


def workload_config(config, B, T, engine, world):
    cfg = {
        'workload': ('configs[2]: parallel WaveNet 4-flow IAF student ({}), batch {}x{} samples '
                     'per GPU, synthetic mel, device-drawn noise' if config == 'student' else
                     'configs[3] per-GPU share: ClariNet Gaussian IAF ({}), batch {}x{} samples per GPU, '
                     'synthetic mel, device-drawn noise').format(CONFIGS[config], B, T),
        'clips_per_gpu': B, 'samples_per_clip': T,
        'l2': 'flushed between timed iterations (256 MB write)', 'parallelism': 'clips x{}'.format(world),
    }
    if engine is not None:
        cfg['engine'] = engine
    return cfg


# ------------------------------------------------------------------------------------------------------------------
# secondary blocks
# ------------------------------------------------------------------------------------------------------------------
def time_steps(fn, steps, warmup, flush, dev, barrier):
    """fn(i) enqueues one step; returns per-step CUDA-event times in ms."""
    import torch
    for i in range(warmup):
        fn(i)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        if flush is not None:
            flush.zero_()
        evs[i][0].record()
        fn(i)
        evs[i][1].record()
    barrier()
    return [a.elapsed_time(b) for a, b in evs]


def fastgen_bench(device, steps, hbm_peak, with_cpu):
    """BASELINE configs[1]: teacher WaveNet (wavenet_mol.json, 30 layers, random init) free-running generation.
    `value` = batch 1 on the latency engine; `batched` = 8 utterances on the batched engine; encoding resident in
    HBM, in-kernel RNG; `e2e` goes through nsw_fastgen_run_host (encoding H2D + audio D2H inside)."""
    import torch
    from nsynth_wavenet_b200 import FastgenEngine
    from nsynth_wavenet_b200.weights_init import init_teacher_weights
    hp = load_hparams('wavenet_mol.json')
    eng = FastgenEngine(hp, init_teacher_weights(hp, seed=12345), device=device)
    g = torch.Generator(device='cpu').manual_seed(1)
    enc_h = (torch.rand((1, steps, 256), generator=g) * 2 - 1)
    enc = enc_h.to('cuda:%d' % device)
    eng.run_device(enc[:, :2048].contiguous(), seed=1)       # warm-up
    torch.cuda.synchronize()
    eng.run_device(enc, seed=2)
    torch.cuda.synchronize()
    ms = eng.last_timing()
    sps = steps / (ms * 1e-3)
    blocks_bytes = 32 * 128 * 9224 * 4          # per-step weight blocks streamed L2 -> shared memory (all CTAs)
    out = {'metric': 'fastgen audio samples/sec (wavenet_mol.json, batch 1)', 'value': sps,
           'unit': UNIT, 'rtf': sps / 16000.0, 'steps': steps, 'ms': ms, 'us_per_step': 1e3 * ms / steps,
           'engine': 'latency (one persistent kernel, tagged-word exchange, no grid barrier)',
           'roofline': {
               'bound': 'hbm', 'unit': 'GB/s', 'peak': hbm_peak,
               # 118.8 MB of weights touched per step is the algorithmic traffic
               'achieved': FASTGEN_ALGORITHMIC_BYTES * sps / 1e9,
               'frac': FASTGEN_ALGORITHMIC_BYTES * sps / 1e9 / hbm_peak,
               # measured DRAM reads of the kernel (profiles/r01/fastgen_ncu_run56.json: 76.7 MB per sample; the rest
               # of the stream is served from the L2-resident evict_last blocks)
               'dram_read_gbs': 76.7e6 * sps / 1e9, 'frac_dram': 76.7e6 * sps / 1e9 / hbm_peak,
               # L2 -> shared-memory traffic (not HBM): 151 MB per step incl. the precomputed M = W2 Wr matrices
               'l2_to_smem_gbs': blocks_bytes * sps / 1e9,
               'note': 'the step is bound by 32 exchange hops (~1.3 us each), not by bandwidth'}}
    # end to end through the C ABI host call: 8000 steps, encoding (8 MB) copied H2D, audio D2H inside
    Te = min(steps, 8000)
    enc_np = enc_h[:, :Te].contiguous().numpy()
    eng.run_host(enc_np, seed=3)                 # same length: the workspace (cond planes) is grown once, like a warm server
    t0 = time.perf_counter()
    eng.run_host(enc_np, seed=4)
    dt = time.perf_counter() - t0
    out['e2e'] = {'value': Te / dt, 'unit': UNIT, 'rtf': Te / dt / 16000.0, 'steps': Te,
                  'h2d_bytes_per_step': 256 * 4, 'd2h_bytes_per_step': 4,
                  'note': 'nsw_fastgen_run_host: pageable encoding H2D + hoisted cond GEMM + kernel + audio D2H'}
    # batched engine: 8 utterances per weight pass
    Bb, Tb = 8, min(steps, 8000)
    encb = (torch.rand((Bb, Tb, 256), generator=g) * 2 - 1).to('cuda:%d' % device)
    eng.run_device(encb[:, :512].contiguous(), seed=5)
    torch.cuda.synchronize()
    eng.run_device(encb, seed=6)
    torch.cuda.synchronize()
    msb = eng.last_timing()
    out['batched'] = {'batch': Bb, 'steps': Tb, 'value': Bb * Tb / (msb * 1e-3), 'unit': UNIT,
                      'rtf_aggregate': Bb * Tb / (msb * 1e-3) / 16000.0, 'us_per_step': 1e3 * msb / Tb,
                      'engine': 'batched (one grid barrier per layer, weights streamed once for all rows)'}
    eng.close()
    # wavenet_ce.json as shipped: gate 1024 (double_gate_width default), mu-law, 256-way softmax head
    try:
        hpc = load_hparams('wavenet_ce.json')
        ce = FastgenEngine(hpc, init_teacher_weights(hpc, seed=12345), device=device)
        ce.run_device(encb[:, :256].contiguous(), seed=7)
        torch.cuda.synchronize()
        Tc = min(Tb, 4000)
        ce.run_device(encb[:, :Tc].contiguous(), seed=8)
        torch.cuda.synchronize()
        msc = ce.last_timing()
        out['ce_double_gate_batched'] = {'batch': Bb, 'steps': Tc, 'value': Bb * Tc / (msc * 1e-3), 'unit': UNIT,
                                         'rtf_aggregate': Bb * Tc / (msc * 1e-3) / 16000.0,
                                         'us_per_step': 1e3 * msc / Tc}
        ce.close()
    except Exception as ex:
        out['ce_double_gate_batched'] = {'error': str(ex)[:200]}
    if with_cpu:
        out['cpu_baseline'] = cpu_fastgen_run(2000)
    return out


def distill_bench(device, tc_peak, world, dev, barrier, max_over_ranks):
    """BASELINE configs[4] forward on every rank at the author's per-GPU batch (7 x 7680: the reference's
    total_batch_size 28 is 4 GPUs x 7; N GPUs here hold 7 N clips, weak scaling): student forward + teacher
    full-sequence forward on the student's output + 100-sample MoL cross-entropy (parallel_wavenet.py:361-402), the
    loss all-reduced over ranks (the forward's only collective)."""
    import torch
    import torch.distributed as dist
    from nsynth_wavenet_b200 import IAFEngine, TeacherEngine
    from nsynth_wavenet_b200.weights_init import init_student_weights, init_teacher_weights
    thp = load_hparams('wavenet_mol.json')
    shp = load_hparams('student')
    st = IAFEngine(shp, init_student_weights(shp, seed=12345), device=device)
    te = TeacherEngine(thp, init_teacher_weights(thp, seed=12345), device=device)
    g = torch.Generator(device='cpu').manual_seed(100 + int(os.environ.get('RANK', '0')))
    mel = torch.rand((7, 39, 80), generator=g).to(dev)
    res = None
    evs, tes = [], []
    barrier()
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = st.forward_device(mel, None, seed=i, quantize=False)
        te_out = te.forward_device(out['x'], mel)
        res = te.mol_score(te_out, out['mean_tot'], out['scale_tot'], out['log_scale_tot'],
                           num_samples=100, seed=i)
        if world > 1:
            lt = torch.tensor([res['kl_loss']], dtype=torch.float64, device=dev)
            dist.all_reduce(lt)
            res = dict(res, kl_loss_mean_over_ranks=float(lt.item()) / world)
        e1.record()
        torch.cuda.synchronize()
        evs.append(e0.elapsed_time(e1))
        tes.append(te.last_timing())
    ms = max_over_ranks(float(np.median(evs[1:])))
    te_ms = float(np.median(tes[1:]))
    n = 7 * 7680
    issued = TEACHER_FLOP_PER_SAMPLE * n * SPLIT_PRODUCTS / (te_ms * 1e-3) / 1e12
    st.close()
    te.close()
    return {'metric': 'configs[4] distillation forward (student + teacher + 100-sample MoL CE), 7x7680 per GPU, '
                      '{} GPU(s): {} clips'.format(world, 7 * world),
            'ms': ms, 'clips_per_s': 7 * world / (ms * 1e-3), 'teacher_forward_ms': te_ms, 'losses': res,
            'roofline_teacher': {
                'kernel': 'conv_gemm_tc2_kernel<EPI_GATE,1> / <EPI_ROWS,0> (teacher layers, cta_group::2)', 'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': tc_peak,
                'achieved': issued, 'frac': issued / tc_peak,
                'achieved_fp32_equivalent': issued / SPLIT_PRODUCTS,
                'note': 'issued fp16 tensor flops = 3 x the algorithmic 67.4 MFLOP per sample (split-fp16 products)'}}


def clarinet_bench(device, rank, world, dev, flush, barrier, max_over_ranks, steps):
    """BASELINE configs[3]: ClariNet Gaussian IAF (parallel_wavenet_gauss.json, four separate deconv stacks),
    8 clips of 7680 samples per GPU (64 x 7680 over 8 GPUs), weights broadcast from rank 0."""
    from nsynth_wavenet_b200 import IAFEngine, parallel
    from nsynth_wavenet_b200.weights_init import init_student_weights
    import torch
    hp = load_hparams('clarinet')
    w0 = init_student_weights(hp, seed=12345)
    if rank != 0:
        w0 = {k: np.zeros_like(v) for k, v in w0.items()}
    eng = IAFEngine(hp, parallel.broadcast_weights(w0, device=dev), device=device)
    g = torch.Generator(device='cpu').manual_seed(200 + rank)
    mel = torch.rand((8, 39, 80), generator=g).to(dev)
    T = eng.length(39)
    out = {k: torch.empty((8, T), dtype=torch.float32, device=dev) for k in ('x', 'mean_tot', 'scale_tot', 'log_scale_tot')}
    ts = time_steps(lambda i: eng.forward_device(mel, None, seed=i, quantize=True, out=out), steps, 3, flush, dev, barrier)
    total = max_over_ranks(float(np.sum(ts)))
    eng.close()
    return {'metric': 'configs[3] ClariNet Gaussian IAF samples/s, 8x7680 per GPU, {} GPU(s)'.format(world),
            'value': world * 8 * T * steps / (total * 1e-3), 'unit': UNIT, 'ms_per_step': total / steps,
            'rtf': world * 8 * T * steps / (total * 1e-3) / 16000.0}


def python_e2e(hp, B, F, T, reps):
    """The call a user of the reference makes: wavenet.parallelgen.synthesis(hparams, mel, save_paths, checkpoint_path)
    with pageable NumPy arrays, the checkpoint on disk and wav files written (parallelgen.py:22-51).  First call =
    checkpoint read + engine build; repeat calls hit the engine cache (checkpoint.cached_engine)."""
    from nsynth_wavenet_b200 import checkpoint as ckpt
    from nsynth_wavenet_b200.weights_init import init_student_weights
    from wavenet import parallelgen
    rng = np.random.default_rng(777)
    mel = rng.uniform(0, 1, (B, F, 80)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        ck = ckpt.save_weights(os.path.join(d, 'model.ckpt-1'), init_student_weights(hp, seed=12345), ema=True)
        paths = [os.path.join(d, 'gen_%d.wav' % i) for i in range(B)]
        t0 = time.perf_counter()
        parallelgen.synthesis(hp, mel, paths, ck, seed=1)
        first = time.perf_counter() - t0
        ts = []
        for i in range(reps):
            t0 = time.perf_counter()
            parallelgen.synthesis(hp, mel, paths, ck, seed=2 + i)
            ts.append(time.perf_counter() - t0)
        ckpt.clear_engine_cache()
    rep = float(np.median(ts))
    return {'value': B * T / rep, 'unit': UNIT, 'repeat_call_ms': 1e3 * rep, 'first_call_s': first,
            'h2d_bytes_per_step': int(B * F * 80 * 4), 'd2h_bytes_per_step': int(B * T * 4),
            'note': 'parallelgen.synthesis end to end: pageable mel H2D, forward, audio D2H, {} float32 wav files '
                    'written; repeat calls reuse the cached engine (keyed on checkpoint path + mtime)'.format(B)}


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
    ap.add_argument('--ref-batch', type=int, default=8, help='clips per step of the CPU arms (same config as the GPU arm)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--sustained-seconds', type=float, default=2.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fastgen', action='store_true')
    ap.add_argument('--no-distill', action='store_true')
    ap.add_argument('--no-clarinet', action='store_true')
    ap.add_argument('--no-sustained', action='store_true')
    ap.add_argument('--no-python-e2e', action='store_true')
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
    from nsynth_wavenet_b200 import IAFEngine, _lib, parallel
    from nsynth_wavenet_b200.weights_init import init_student_weights

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        return parallel.max_over_ranks(v, device=dev)

    hp = load_hparams(args.config)
    # ---- weights: rank 0 initialises, NCCL broadcasts one flat blob (the only collective) ----
    w0 = init_student_weights(hp, seed=12345)
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
    clocks = sampler.finish()
    gpu_ms = sum(a.elapsed_time(b) for a, b in evs)
    total_ms = max_over_ranks(gpu_ms)
    ms_per_step = total_ms / args.steps
    value = world * B * T * args.steps / (total_ms * 1e-3)

    # ---- sustained: the same step (same L2 flush between steps) looped for >= 2 s of wall clock ----
    sustained = None
    if not args.no_sustained:
        s2 = ClockSampler(local_rank)
        s2.start()
        per = []
        t_start = time.perf_counter()
        block = 0
        while time.perf_counter() - t_start < args.sustained_seconds:
            per += time_steps(step, 100, 0, flush, dev, lambda: torch.cuda.synchronize())
            block += 1
        wall = time.perf_counter() - t_start
        sc = s2.finish()
        med = max_over_ranks(float(np.median(per)))
        sustained = {'ms_per_step_median': med, 'ms_per_step_mean': float(np.mean(per)),
                     'ms_per_step_p95': float(np.percentile(per, 95)), 'steps': len(per), 'wall_s': wall,
                     'value': world * B * T / (med * 1e-3), 'unit': UNIT, 'clocks': sc}

    # ---- e2e: the reference-facing host call, pinned host buffers, H2D + D2H inside ----
    x_h = np.empty((B, T), np.float32)
    x_pin = torch.from_numpy(x_h).pin_memory()
    mel_np = np.array(mel_h.numpy())       # pageable copy
    for i in range(2):
        eng.forward_host(mel_np, None, seed=i, quantize=True)
    barrier()
    e2e_s = 0.0
    for i in range(args.steps):
        flush.zero_()                      # same L2 hygiene as the device-timed loop (not timed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()           # the call is synchronous: returns with x in host memory
        _lib.check(lib.nsw_iaf_forward_host(eng._h, mel_h.data_ptr(), None, 2000 + i, B, F, 1,
                                            x_pin.data_ptr(), None, None, None, None))
        e2e_s += time.perf_counter() - t0
    e2e_value = world * B * T * args.steps / max_over_ranks(e2e_s)
    pg_s = 0.0
    for i in range(args.steps):            # the same call with pageable NumPy arrays (what Python callers hold)
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.forward_host(mel_np, None, seed=3000 + i, quantize=True)
        pg_s += time.perf_counter() - t0
    e2e_pageable = world * B * T * args.steps / max_over_ranks(pg_s)

    # ---- rooflines of the dominant kernels, measured live with CUDA events on the launch stream inside the
    #      library (separate profiled passes) ----
    eng.set_profiling(True)
    prof = []
    for i in range(5):
        flush.zero_()
        step(i)
        torch.cuda.synchronize()
        prof.append(eng.last_timing())
    eng.set_profiling(False)
    stage = {k: float(np.median([p[k] for p in prof])) for k in prof[0]}
    n_layers = int(sum(hp.num_iaf_layers))
    n_flows = len(hp.num_iaf_layers)
    n_planes = n_layers + n_flows
    hbm_peak, tc_peak, peak_kind = measured_peaks()
    fused = eng.engine == 'tc3'
    # algorithmic bytes of the 'layers' stage: 768 B per (sample, layer); engine tc3 also runs the
    # start conv (read x 4 B + write l 256 B) and the head (l 256 + cond 256 + ~40 B of x / totals) of every
    # flow inside the same kernel, so they are part of both the bytes and the time
    stage_bytes = B * T * (LAYER_BYTES_PER_SAMPLE * n_layers + (n_flows * (260 + 552) if fused else 0))
    n_launches = n_flows if fused else n_layers
    layer_gbs = stage_bytes / (stage['layers'] * 1e-3) / 1e9
    # issued tensor-core flops of the same stage: 3 split-fp16 products per algorithmic MAC
    stage_flop = B * T * (LAYER_FLOP_PER_SAMPLE * n_layers + (HEAD_FLOP_PER_SAMPLE * n_flows if fused else 0))
    issued_tflops = stage_flop * SPLIT_PRODUCTS / (stage['layers'] * 1e-3) / 1e12
    pair_flow = eng.engine == 'tc3' and os.environ.get('NSW_FLOW_PAIR', '1') != '0' and T % 256 == 0
    kernel_name = {'tc3': 'iaf_flow_pair_kernel (cta_group::2)' if pair_flow else 'iaf_flow_tc_kernel',
                   'tc2': 'iaf_layer_tc_kernel'}.get(eng.engine, 'iaf_layer_kernel')
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (one 10-layer flow launch);
    # only valid for the shape and engine it was captured on
    traffic, traffic_src, traffic_alg, frac_dram, ncu_extra = None, None, None, None, None
    cap, cap_src = (None, None)
    if fused and (B, T) == (8, 7680):
        cap, cap_src = ncu_capture('iaf_flow_pair' if pair_flow else 'iaf_flow_tc')
        if cap is None:
            cap, cap_src = ncu_capture('iaf_flow_')
    if cap:
        rd, wr = metric_value(cap, 'dram__bytes_read.sum'), metric_value(cap, 'dram__bytes_write.sum')
        scale = 1e6 if rd is not None and rd < 1e5 else 1.0     # r01 summaries are in MB
        traffic = (rd + wr) * scale
        traffic_alg = B * T * (LAYER_BYTES_PER_SAMPLE * 10 + 812)
        traffic_src = cap_src + ': dram__bytes_read+write of one flow-kernel launch (10-layer flow incl. start conv and head)'
        # the four flows move 60 / 10 x this; divide by the live stage time
        frac_dram = traffic * (n_layers / 10.0) / (stage['layers'] * 1e-3) / 1e9 / hbm_peak
        ncu_extra = {k: metric_value(cap, k) for k in cap
                     if k.startswith(('sm__pipe_tensor', 'l1tex__data_pipe', 'smsp__issue_active', 'sm__inst_executed_pipe',
                                      'l1tex__data_bank', 'smsp__average_warp'))}
    cond_tflops = COND_FLOP_PER_SAMPLE_PLANE * n_planes * B * T / (stage['cond'] * 1e-3) / 1e12

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.config, B, T, eng.engine, world),
        'rtf': value / 16000.0,
        'rtf_per_gpu': value / 16000.0 / world,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(B * F * 80 * 4),
                'd2h_bytes_per_step': int(B * T * 4), 'value_pageable_buffers': e2e_pageable},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'sustained': sustained,
        'roofline': {
            'kernel': kernel_name,
            # what bounds the kernel: the tensor pipe's operand fetch (N=64 SS MMAs are shared-memory-bandwidth
            # bound at 49 instead of 32 cycles; profiles/r02 ncu: shared-memory data pipe vs DRAM), NOT HBM
            'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': tc_peak,
            'achieved': issued_tflops, 'frac': issued_tflops / tc_peak, 'frac_tensor': issued_tflops / tc_peak,
            # the same against the SUSTAINED dense peak (MEASURED_PEAKS.json: a 4 s GEMM under the power cap)
            'frac_tensor_vs_sustained_peak': issued_tflops / sustained_tc_peak() if sustained_tc_peak() else None,
            'achieved_fp32_equivalent': issued_tflops / SPLIT_PRODUCTS,
            # SURVEY's per-layer-kernel byte model (768 B per sample and layer) against the HBM peak: a model
            # throughput, kept for continuity with round 1 -- the fused kernel never moves those bytes
            'frac_model_hbm': layer_gbs / hbm_peak, 'model_gbs': layer_gbs, 'hbm_peak_gbs': hbm_peak,
            # real DRAM traffic (ncu bytes of the committed capture, scaled to the 60 layers) / live stage time / peak
            'frac_dram': frac_dram,
            'traffic': traffic, 'traffic_launch_algorithmic_bytes': traffic_alg, 'traffic_source': traffic_src,
            'ncu': ncu_extra,
            'unit_of_launch': ('one flow (start conv + all residual layers + head) over the whole batch; '
                               'flops_per_launch and launch_ms are the averages over the flows'
                               if fused else 'one residual layer over the whole batch'),
            'peak_source': peak_kind + ' (MEASURED_PEAKS.json bf16_tflops, burst; hbm_gbs for the *_hbm / *_dram fractions)',
            'flops_per_launch_issued': stage_flop * SPLIT_PRODUCTS / n_launches, 'bytes_per_launch_model': stage_bytes / n_launches,
            'launch_ms': stage['layers'] / n_launches,
            'share_of_step': stage['layers'] / stage['total'],
        },
        'roofline_cond_gemm': {
            'kernel': 'cond_proj_tc2_kernel<256> (cta_group::2)' if eng.engine == 'tc3' else 'conv_gemm (cond projections)', 'bound': 'tensor',
            'achieved': cond_tflops * SPLIT_PRODUCTS, 'peak': tc_peak, 'unit': 'TFLOP/s',
            'frac': cond_tflops * SPLIT_PRODUCTS / tc_peak, 'achieved_fp32_equivalent': cond_tflops,
            'note': 'issued fp16 tensor flops = 3 x the algorithmic 2*256*64 per (sample, plane)',
        },
        'stage_ms': stage,
    }

    # ---- the call a Python user makes: parallelgen.synthesis with a checkpoint on disk (rank 0) ----
    if rank == 0 and not args.no_python_e2e:
        try:
            line['e2e_python'] = python_e2e(hp, B, F, T, 10)
        except Exception as ex:
            line['e2e_python'] = {'error': str(ex)[:300]}

    # ---- configs[3] and configs[4] at this N (max over ranks) ----
    if not args.no_clarinet and args.config == 'student':
        try:
            cl = clarinet_bench(local_rank, rank, world, dev, flush, barrier, max_over_ranks, max(5, min(args.steps, 20)))
            if rank == 0:
                line['clarinet'] = cl
        except Exception as ex:
            if rank == 0:
                line['clarinet'] = {'error': str(ex)[:300]}
    if not args.no_distill:
        try:
            ds = distill_bench(local_rank, tc_peak, world, dev, barrier, max_over_ranks)
            if rank == 0:
                line['distill'] = ds
        except Exception as ex:  # reported, never silently dropped
            if rank == 0:
                line['distill'] = {'error': str(ex)[:300]}

    # ---- secondary metric: autoregressive fastgen (BASELINE configs[1]), rank 0 ----
    if rank == 0 and not args.no_fastgen:
        try:
            line['fastgen'] = fastgen_bench(local_rank, args.fastgen_steps, hbm_peak,
                                            with_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as ex:
            line['fastgen'] = {'error': str(ex)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample (about cpu-seconds of work) of the SAME workload: 8 x 7680 per step
        v1, ms1, cores, _ = cpu_reference_run(hp, args.ref_batch, F, 1, 1)
        reps = int(max(1, min(20, args.cpu_seconds / (ms1 * 1e-3))))
        v, ms, cores, sample = cpu_reference_run(hp, args.ref_batch, F, reps, 0)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': sample}
    elif rank == 0:
        line['cpu_baseline'] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
