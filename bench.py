"""Benchmark of the frame-based feature hot path (BASELINE.json metric:
"MFCC frames/sec at 1/2/4/8 B200; achieved HBM GB/s vs roofline").

    python bench.py --gpus N --steps K --warmup W        (torchrun for N > 1)
    python bench.py --impl reference ...                  (CPU arm)

Workload (config.workload): BASELINE.json configs[2], the MFCC configuration
the metric is quoted on that fits one GPU -- MfccProcessor (reference
defaults, dither=1.0) + per-utterance CMVN + DeltaPostProcessor(order=2) on
10 000 synthetic 16 kHz 10 s utterances PER GPU (weak scaling: utterances are
independent, each rank owns its shard, no data-path collective).  One "step"
is one pass of the whole pipeline over the rank's batch.

* value      : frames/s with the int16 PCM already resident in HBM (CUDA
               events around K steps, max over ranks, whole-job aggregate)
* e2e        : same metric through the host API with pinned HOST buffers:
               H2D of the PCM and D2H of the features inside the timed region
* roofline   : dominant kernel (fused_features_512_kernel): algorithmic
               bytes/launch (372 B/frame: 320 B int16 PCM read + 13 floats
               written) / its mean duration, vs the measured HBM peak
* cpu_baseline: the C oracle port of the reference path (Kaldi restatement),
               OpenMP over utterances on all host cores, bounded sample
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 16000
UTT_SAMPLES = 160000            # 10 s
FRAMES_PER_UTT = 998
BYTES_PER_FRAME_KERNEL = 320 + 4 * 13     # PCM read once + 13 cepstra written
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fused_features_512_kernel
# launch over 9.98e6 frames (ncu capture of `bench.py --steps 2 --warmup 2`,
# profiles/r01_final_ncu_traffic_fused_features_512.csv): 3 242 790 912 +
# 512 223 488 B = 376.3 B/frame, i.e. 1.011 x the algorithmic bytes (the
# extra 2 B/frame is the 32-byte tile descriptor per 16 frames)
NCU_TRAFFIC_BYTES_PER_FRAME = (3242790912 + 512223488) / 9980000.0
TRAFFIC_SOURCE = ('ncu capture of this launch shape, '
                  'profiles/r01_final_ncu_traffic_fused_features_512.csv')
# what actually bounds the dominant kernel (ncu --set full of the same kernel on
# 2 000 utterances, profiles/r01_final_ncu_fused_features_512_dither*.txt)
NCU_LIMITS = {
    'source': 'profiles/r01_final_ncu_fused_features_512_dither1.0.txt',
    'issue_slots_busy_pct': 72.9, 'smem_data_pipe_busy_pct': 64.0,
    'warp_instructions_per_frame': 936, 'smem_wavefronts_per_frame': 215,
    'dram_throughput_pct': 4.0}
BYTES_PER_FRAME_PIPELINE = 320 + 4 * 39   # + delta/cmvn output
FLOPS_PER_FRAME = 17000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--utts', type=int, default=10000,
                    help='utterances per GPU')
    ap.add_argument('--dither', type=float, default=1.0)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--chunk-utts', type=int, default=512,
                    help='utterances per chunk of the host pipeline (e2e)')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def synth_pcm_device(nutts, rank, torch):
    """Synthetic corpus of BASELINE.md section 3 generated on the device:
    5-harmonic tone (f0 ~ U(80, 300) Hz, amplitude 3000/h) + N(0, 500^2)
    noise, clipped to int16; seeded by (20260925, rank)."""
    gen = torch.Generator(device='cuda')
    gen.manual_seed(20260925 + 7919 * rank)
    out = torch.empty((nutts, UTT_SAMPLES), dtype=torch.int16, device='cuda')
    t = torch.arange(UTT_SAMPLES, device='cuda', dtype=torch.float32) / SAMPLE_RATE
    step = 250
    for b in range(0, nutts, step):
        n = min(step, nutts - b)
        f0 = 80 + 220 * torch.rand((n, 1), generator=gen, device='cuda')
        x = 500.0 * torch.randn((n, UTT_SAMPLES), generator=gen, device='cuda')
        for h in range(1, 6):
            ph = 2 * np.pi * torch.rand((n, 1), generator=gen, device='cuda')
            x += (3000.0 / h) * torch.sin(2 * np.pi * h * f0 * t[None, :] + ph)
        out[b:b + n] = x.round().clamp_(-32768, 32767).to(torch.int16)
    return out.reshape(-1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,'
             'clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index),
                 '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        if not sm and self.rows:   # region shorter than the sampling period
            parts = [p.strip() for p in self.rows[-1][1].split(',')]
            try:
                sm, mx = [float(parts[0])], [float(parts[1])]
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def oracle_kwargs(dither):
    return dict(dither=dither)


def cpu_baseline(pcm_host, nutts_avail, dither, budget_s=12.0):
    """Times the CPU oracle port (mfcc + cmvn + delta, OpenMP over utterances)
    on a bounded sample of the same workload"""
    import oracle
    cores = os.cpu_count() or 1
    offs = np.arange(nutts_avail + 1, dtype=np.int64) * UTT_SAMPLES

    def run(n):
        t0 = time.perf_counter()
        out, fofs = oracle.pipeline_batch(
            'mfcc', pcm_host[:n * UTT_SAMPLES], offs[:n + 1], cmvn=True,
            norm_vars=True, delta_order=2, delta_window=2, nthreads=cores,
            dither=dither)
        return time.perf_counter() - t0, int(fofs[-1])
    probe = min(nutts_avail, max(cores, 8))
    dt, frames = run(probe)
    rate = frames / dt
    n = int(min(nutts_avail, max(probe, rate * budget_s / FRAMES_PER_UTT)))
    # passes over the sample until about budget_s of CPU work has been timed
    reps = int(max(1, min(64, round(budget_s * rate / (n * FRAMES_PER_UTT)))))
    dt, frames = 0.0, 0
    for _ in range(reps):
        d, f = run(n)
        dt += d
        frames += f
    return {'value': frames / dt, 'unit': 'frames/s', 'cores': cores,
            'kind': 'port',
            'sample': f'{reps} pass(es) over {n} of the synthetic 10 s '
                      f'utterances ({frames} frames) in {dt:.2f} s, C oracle, '
                      f'OpenMP over utterances'}, n, dt


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()

    config = {
        'workload': ('BASELINE configs[2]: MfccProcessor (13 ceps, 23 mel, '
                     'reference defaults) + CMVN per utterance (norm_vars) + '
                     'DeltaPostProcessor(order=2, window=2) on '
                     f'{args.utts} synthetic 16 kHz 10 s utterances per GPU'),
        'utterances_per_gpu': args.utts, 'frames_per_utterance': FRAMES_PER_UTT,
        'dither': args.dither, 'sharding': f'utterances x{world} (no collective)',
        'l2': 'inputs (3.2 GB/GPU) larger than the 126 MB L2',
    }

    if args.impl == 'reference':
        # the reference's own CPU implementation cannot be imported (pykaldi is
        # absent): the arm times the oracle port on all host cores
        if rank != 0:
            return
        import torch
        nutts = min(args.utts, 2048)
        rng = np.random.default_rng(20260925)
        # same distribution as the GPU corpus, generated on the host
        t = np.arange(UTT_SAMPLES, dtype=np.float32) / SAMPLE_RATE
        pcm = np.empty(nutts * UTT_SAMPLES, dtype=np.int16)
        for u in range(nutts):
            f0 = rng.uniform(80, 300)
            x = 500.0 * rng.standard_normal(UTT_SAMPLES, dtype=np.float32)
            for h in range(1, 6):
                x += (3000.0 / h) * np.sin(
                    2 * np.pi * h * f0 * t + rng.uniform(0, 2 * np.pi))
            pcm[u * UTT_SAMPLES:(u + 1) * UTT_SAMPLES] = np.clip(
                np.round(x), -32768, 32767)
        values = []
        for step in range(args.warmup + args.steps):
            base, n, dt = cpu_baseline(pcm, nutts, args.dither, budget_s=8.0)
            if step >= args.warmup:
                values.append((base, n, dt))
        value = float(np.mean([b['value'] for b, _, _ in values]))
        base = values[-1][0]
        base['value'] = value
        print(json.dumps({
            'impl': 'reference', 'metric': 'MFCC frames/sec', 'value': value,
            'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup,
            'ms_per_step': float(np.mean([dt for _, _, dt in values]) * 1e3),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'cpu_baseline': base,
            'e2e': {'value': value, 'unit': 'frames/s',
                    'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ['NCCL_DEBUG'] = 'WARN'      # keep stdout to the one JSON line
        dist.init_process_group(
            'nccl', device_id=torch.device('cuda', local_rank))
    from shennong_b200 import _lib, engine
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.postprocessor import DeltaPostProcessor
    from shennong_b200.processor import MfccProcessor

    proc = MfccProcessor(dither=args.dither)
    pipe = FusedPipeline(proc, delta=DeltaPostProcessor(order=2, window=2),
                         cmvn='utterance', norm_vars=True)
    plans = pipe._plans()
    nutts = args.utts
    pcm_dev = synth_pcm_device(nutts, rank, torch)
    pad = torch.zeros(64, dtype=torch.int16, device='cuda')
    pcm_dev = torch.cat([pcm_dev, pad])
    starts = np.arange(nutts, dtype=np.int64) * UTT_SAMPLES
    lengths = np.full(nutts, UTT_SAMPLES, dtype=np.int64)
    packed = engine.PackedAudio.from_packed(None, starts, lengths, dev=pcm_dev)
    batch = engine.Batch(plans['feat'], packed)
    layout = engine.RowLayout(batch=batch)
    total_frames = batch.total_frames
    base = torch.empty((total_frames, 13), dtype=torch.float32, device='cuda')
    out = torch.empty((total_frames, 39), dtype=torch.float32, device='cuda')
    L = _lib.lib()

    ev_feat = [(torch.cuda.Event(enable_timing=True),
                torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]

    def step(i=None, seed=1):
        if i is not None:
            ev_feat[i][0].record()
        engine.compute_features(plans['feat'], batch, seed=seed, out=base)
        if i is not None:
            ev_feat[i][1].record()
        stats = engine.cmvn_accumulate(base, layout)
        norm = engine.cmvn_norm(stats, True, False)
        engine.deltas(base, layout, 2, 2, norm=norm, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(seed=100 + w)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = L.snb_launch_count()
    barrier()
    t0 = time.time()
    e0, e1 = (torch.cuda.Event(enable_timing=True),
              torch.cuda.Event(enable_timing=True))
    e0.record()
    for i in range(args.steps):
        step(i, seed=1000 + i)
    e1.record()
    barrier()
    t1 = time.time()
    launches = L.snb_launch_count() - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    ms_feat = float(np.mean([a.elapsed_time(b) for a, b in ev_feat]))
    times = torch.tensor([ms_total, ms_feat], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_feat = float(times[0]), float(times[1])
    ms_per_step = ms_total / args.steps
    value = world * total_frames / (ms_per_step * 1e-3)

    # ---- end-to-end: pinned host PCM in, pinned host features out ----------
    e2e = None
    if not args.no_e2e:
        host_pcm = torch.empty(pcm_dev.numel(), dtype=torch.int16,
                               pin_memory=True)
        host_pcm.copy_(pcm_dev)
        out_host = torch.empty((total_frames, 39), dtype=torch.float32,
                               pin_memory=True)
        torch.cuda.synchronize()
        # raw PCIe ceilings of this box (plain pinned copies, one direction)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scratch = torch.empty_like(pcm_dev)
        c0.record(); scratch.copy_(host_pcm, non_blocking=True); c1.record()
        torch.cuda.synchronize()
        h2d_gbs = host_pcm.numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        c0.record(); out_host.copy_(out, non_blocking=True); c1.record()
        torch.cuda.synchronize()
        d2h_gbs = out.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del scratch
        nrep = max(2, min(args.steps, 5))
        for _ in range(2):                                           # warm-up
            pipe.run_host(host_pcm, starts, lengths, out_host=out_host,
                          chunk_utts=args.chunk_utts)
        barrier()
        rep_ms = []
        t_e0 = time.perf_counter()
        for _ in range(nrep):
            t_r = time.perf_counter()
            pipe.run_host(host_pcm, starts, lengths, out_host=out_host,
                          chunk_utts=args.chunk_utts)
            rep_ms.append((time.perf_counter() - t_r) * 1e3)
        barrier()
        dt = (time.perf_counter() - t_e0) / nrep
        tt = torch.tensor([dt], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        e2e = {'value': world * total_frames / dt, 'unit': 'frames/s',
               'h2d_bytes_per_step': int(nutts * UTT_SAMPLES * 2),
               'd2h_bytes_per_step': int(total_frames * 39 * 4),
               'ms_per_step': dt * 1e3,
               'ms_each_step': [round(t, 2) for t in rep_ms],
               'api': 'FusedPipeline.run_host (chunked H2D/compute/D2H, '
                      f'{args.chunk_utts} utterances per chunk)',
               'pcie_h2d_gbs': h2d_gbs, 'pcie_d2h_gbs': d2h_gbs}

    # ---- collection (outside the step): NCCL all-gather of the feature blocks --
    gather = None
    if world > 1:
        from shennong_b200.distributed import gather_rows
        full, _ = gather_rows(out)                                   # warm-up
        del full
        barrier()
        g0, g1 = (torch.cuda.Event(enable_timing=True),
                  torch.cuda.Event(enable_timing=True))
        g0.record()
        for _ in range(3):
            full, _ = gather_rows(out)
            del full
        g1.record()
        barrier()
        gt = torch.tensor([g0.elapsed_time(g1) / 3], device='cuda',
                          dtype=torch.float64)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        nbytes = int(out.numel() * 4)
        gather = {'ms': float(gt[0]), 'bytes_per_rank': nbytes,
                  'recv_gbs_per_rank': (world - 1) * nbytes / (float(gt[0]) * 1e-3) / 1e9,
                  'api': 'distributed.gather_rows (row counts + one NCCL '
                         'all-gather of the [frames, 39] blocks; not part of '
                         'the timed step)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    feat_gbs = total_frames * BYTES_PER_FRAME_KERNEL / (ms_feat * 1e-3) / 1e9
    roofline = {
        'bound': 'hbm', 'kernel': 'fused_features_512_kernel',
        'achieved': feat_gbs, 'peak': peak, 'unit': 'GB/s',
        'frac': feat_gbs / peak,
        'traffic': int(total_frames * NCU_TRAFFIC_BYTES_PER_FRAME),
        'traffic_source': TRAFFIC_SOURCE,
        'peak_source': peak_src,
        'algorithmic_bytes_per_launch': int(total_frames * BYTES_PER_FRAME_KERNEL),
        'kernel_ms': ms_feat,
        'note': ('the chain is instruction-issue / shared-memory bound (~17 '
                 'kflop/frame, AI ~45 flop/B), not HBM bound: see limits and '
                 'fp32_tflops'),
        'limits': NCU_LIMITS,
        'fp32_tflops': total_frames * FLOPS_PER_FRAME / (ms_feat * 1e-3) / 1e12,
        'pipeline_gbs': total_frames * BYTES_PER_FRAME_PIPELINE
        / (ms_per_step * 1e-3) / 1e9,
    }
    cpu = None
    if not args.no_cpu:
        sample_utts = min(nutts, 2048)
        host_sample = pcm_dev[:sample_utts * UTT_SAMPLES].cpu().numpy()
        cpu, _, _ = cpu_baseline(host_sample, sample_utts, args.dither)

    result = {
        'metric': 'MFCC frames/sec', 'value': value, 'unit': 'frames/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': config, 'clocks': clocks,
        'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
        'cpu_baseline': cpu,
    }
    if gather is not None:
        result['gather'] = gather
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
