"""Benchmark of the frame-based feature hot path (BASELINE.json metric:
"MFCC frames/sec at 1/2/4/8 B200; achieved HBM GB/s vs roofline").

    python bench.py --gpus N --steps K --warmup W        (torchrun for N > 1)
    python bench.py --impl reference ...                  (CPU arm)
    python bench.py --config {1,2,3,4} ...                (BASELINE.json configs[k])

Default workload (config.workload): BASELINE.json configs[2], the MFCC
configuration the metric is quoted on that fits one GPU -- MfccProcessor
(reference defaults, dither=1.0) + per-utterance CMVN + DeltaPostProcessor
(order=2) on 10 000 synthetic 16 kHz 10 s utterances PER GPU (weak scaling:
utterances are independent, each rank owns its shard).  One "step" is one pass
of the whole pipeline over the rank's batch.  configs[3] (PLP + Kaldi pitch,
50 000 utterances over 8 GPUs = 6 250 per GPU) and configs[4] (filterbank +
pitch + delta + CMVN by speaker with VAD, 100 h = 36 000 utterances over 8
GPUs = 4 500 per GPU) are the two 8-GPU configurations of BASELINE.json.

* value      : frames/s with the int16 PCM already resident in HBM (CUDA
               events around K steps, max over ranks, whole-job aggregate).
               For N > 1 the COLLECTION is inside the step: the rank's batch
               is cut in chunks and the rows of chunk k are all-gathered to
               every rank (NCCL over NVLink, own stream) while chunk k + 1 is
               computed; the step ends when every rank holds every row.
* e2e        : same metric through the host API with pinned HOST buffers:
               H2D of the PCM and D2H of the features inside the timed region
               (FusedPipeline.run_host = shennong_b200.stream.StreamRunner)
* e2e_api    : the reference-facing calls on WAV files:
               MfccProcessor().process_all(Utterances) and
               pipeline.extract_features(config, Utterances) (sharded over the
               ranks, rows gathered to every rank)
* roofline   : dominant kernel (fused_features_512_kernel): algorithmic
               bytes/launch (320 B int16 PCM read + 4 B x base columns
               written per frame) / its mean duration, vs the measured HBM peak
* cpu_baseline: the C oracle port of the reference path (Kaldi restatement),
               OpenMP over utterances on all host cores, bounded sample
* check      : rank 0 compares rows of every rank (dither 0) with the oracle
"""

import argparse
import concurrent.futures
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 16000
UTT_SAMPLES = 160000            # 10 s
FRAMES_PER_UTT = 998
UTTS_PER_SPEAKER = 100
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fused_features_512_kernel
# launch over 9.98e6 frames of MFCC-13 (ncu capture of `bench.py --steps 2
# --warmup 2`, profiles/r01_final_ncu_traffic_fused_features_512.csv):
# 3 242 790 912 + 512 223 488 B = 376.3 B/frame, i.e. 1.011 x the algorithmic
# bytes (the extra is the 32-byte tile descriptor per 16 frames)
NCU_TRAFFIC_BYTES_PER_FRAME = (3242790912 + 512223488) / 9980000.0
TRAFFIC_SOURCE = ('ncu capture of this launch shape, '
                  'profiles/r01_final_ncu_traffic_fused_features_512.csv')
FLOPS_PER_FRAME = 17000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=[1, 2, 3, 4],
                    help='index in BASELINE.json configs')
    ap.add_argument('--utts', type=int, default=0,
                    help='utterances per GPU (default: the configuration\'s)')
    ap.add_argument('--dither', type=float, default=1.0)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-api', action='store_true')
    ap.add_argument('--api-utts', type=int, default=4000)
    ap.add_argument('--chunk-utts', type=int, default=0,
                    help='utterances per chunk of the host pipeline (e2e); '
                    '0: the runner\'s default (512, or one round of the '
                    'pitch tracker when the pipeline has pitch)')
    ap.add_argument('--gather-chunks', type=int, default=8,
                    help='chunks of the device-resident step for N > 1')
    ap.add_argument('--gather', default='auto',
                    choices=['auto', 'ce', 'bulk', 'stores', 'nccl', 'none'],
                    help='collection inside the step for N > 1: pushes into '
                    'the peers\' result buffers over NVLink (CUDA IPC) by the '
                    'copy engines (ce), by one-warp CTAs driving the TMA unit '
                    '(bulk), by plain 16-byte stores (stores); or NCCL '
                    'all-gather (nccl); auto: the fastest of ce / bulk / '
                    'nccl in a few untimed trial steps before the warm-up')
    ap.add_argument('--gather-ctas', type=int, default=0,
                    help='CTAs of the bulk / stores kernels (0: 148 / 74 per '
                    'peer, at most 296)')
    ap.add_argument('--gather-base-chunks', type=int, default=-1,
                    help='chunks that travel as base rows + normalisation '
                    'table, the receivers redoing normalise + delta (peer '
                    'modes, pipelines with deltas); -1: 5 of 8 chunks with 6 '
                    'or more peers (measured optimum under the NVLink load '
                    'of 8 ranks, profiles/r02_n2_collection_variants_c.txt), '
                    'else none')
    ap.add_argument('--gather-fanout', type=int, default=1,
                    help='link-load experiment: every push is delivered this '
                    'many times to each peer (N = 2 with 7 carries the NVLink '
                    'traffic per GPU of N = 8); `value` then counts the frames '
                    'of the real ranks only')
    ap.add_argument('--force-chunks', action='store_true',
                    help='cut the step in --gather-chunks chunks even without '
                    'collection (diagnosis of the chunking cost)')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


# --------------------------------------------------------------------------
# configurations (BASELINE.json configs[k])
# --------------------------------------------------------------------------
def make_config(index, dither):
    """(FusedPipeline, description dict) for BASELINE.json configs[index]"""
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.postprocessor import (
        DeltaPostProcessor, VadPostProcessor)
    from shennong_b200.processor import (
        EnergyProcessor, FilterbankProcessor, KaldiPitchPostProcessor,
        KaldiPitchProcessor, MfccProcessor, PlpProcessor)
    if index == 1:
        pipe = FusedPipeline(FilterbankProcessor(num_bins=40, dither=dither))
        info = dict(utts=1000, oracle=('filterbank', dict(num_bins=40)),
                    metric='filterbank-40 frames/sec',
                    workload='BASELINE configs[1]: FilterbankProcessor 40-mel')
    elif index == 2:
        pipe = FusedPipeline(
            MfccProcessor(dither=dither),
            delta=DeltaPostProcessor(order=2, window=2), cmvn='utterance',
            norm_vars=True)
        info = dict(utts=10000, oracle=('mfcc', {}),
                    metric='MFCC frames/sec',
                    workload='BASELINE configs[2]: MfccProcessor (13 ceps, '
                    '23 mel, reference defaults) + CMVN per utterance '
                    '(norm_vars) + DeltaPostProcessor(order=2, window=2)')
    elif index == 3:
        pipe = FusedPipeline(
            PlpProcessor(dither=dither),
            pitch=(KaldiPitchProcessor(), KaldiPitchPostProcessor()))
        info = dict(utts=6250, oracle=('plp', {}),
                    metric='PLP + Kaldi pitch frames/sec',
                    workload='BASELINE configs[3]: PlpProcessor + '
                    'KaldiPitchProcessor (+ post-processing) pipeline, '
                    '50 000 utterances over 8 GPUs')
    else:
        pipe = FusedPipeline(
            FilterbankProcessor(dither=dither),
            delta=DeltaPostProcessor(order=2, window=2), cmvn='speaker',
            vad=VadPostProcessor(), energy=EnergyProcessor(),
            pitch=(KaldiPitchProcessor(), KaldiPitchPostProcessor()))
        info = dict(utts=4500, oracle=('filterbank', {}),
                    metric='full pipeline frames/sec',
                    workload='BASELINE configs[4]: speech-features full '
                    'pipeline (filterbank + Kaldi pitch + delta + CMVN by '
                    'speaker with VAD), 100 h = 36 000 utterances over 8 GPUs')
    return pipe, info


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# --------------------------------------------------------------------------
# synthetic corpus
# --------------------------------------------------------------------------
def synth_utterance(index, nsamples=UTT_SAMPLES):
    """Utterance `index` of the synthetic corpus on the host: 5-harmonic tone
    (f0 ~ U(80, 300) Hz, amplitude 3000/h) + N(0, 500^2) noise, clipped to
    int16; seeded per utterance (same generator as tests/conftest.py)"""
    rng = np.random.default_rng(20260925 + index)
    f0 = rng.uniform(80, 300)
    phases = rng.uniform(0, 2 * np.pi, 5)
    t = np.arange(nsamples) / SAMPLE_RATE
    x = sum(3000.0 / h * np.sin(2 * np.pi * h * f0 * t + phases[h - 1])
            for h in range(1, 6))
    x = x + 500.0 * rng.standard_normal(nsamples)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def synth_host(first, count):
    """Utterances [first, first + count) of the corpus, generated on all host
    cores, as one int16 array [count * UTT_SAMPLES]"""
    out = np.empty(count * UTT_SAMPLES, dtype=np.int16)
    workers = min(os.cpu_count() or 1, 32, max(count, 1))

    def fill(u):
        out[u * UTT_SAMPLES:(u + 1) * UTT_SAMPLES] = synth_utterance(first + u)
    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        list(pool.map(fill, range(count)))
    return out


def synth_pcm_device(nutts, rank, torch):
    """The same distribution generated on the device (bulk of the corpus):
    seeded by (20260925, rank)"""
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


# --------------------------------------------------------------------------
# CPU arm: the oracle port of the same pipeline
# --------------------------------------------------------------------------
def oracle_pass(index, pcm_host, n, dither, cores):
    """One pass of the configuration's pipeline over the first n utterances
    on the CPU oracle; returns (seconds, frames)"""
    import oracle
    offs = np.arange(n + 1, dtype=np.int64) * UTT_SAMPLES
    pcm = pcm_host[:n * UTT_SAMPLES]
    t0 = time.perf_counter()
    if index == 2:
        _, fofs = oracle.pipeline_batch(
            'mfcc', pcm, offs, cmvn=True, norm_vars=True, delta_order=2,
            delta_window=2, nthreads=cores, dither=dither)
        frames = int(fofs[-1])
    elif index == 1:
        _, fofs = oracle.features_batch(
            'filterbank', pcm, offs, nthreads=cores, dither=dither,
            num_bins=40)
        frames = int(fofs[-1])
    else:
        frames = oracle_pass_pitch(index, pcm, n, dither, cores)
    return time.perf_counter() - t0, frames


def oracle_pass_pitch(index, pcm, n, dither, cores):
    """configs[3] / configs[4] on the oracle: the batch entry points for the
    spectral features, a thread pool over utterances for pitch (the C calls
    release the GIL) and the post-processors"""
    import oracle
    offs = np.arange(n + 1, dtype=np.int64) * UTT_SAMPLES
    kind = 'plp' if index == 3 else 'filterbank'
    base, fofs = oracle.features_batch(
        kind, pcm, offs, nthreads=cores, dither=dither)

    def one(u):
        sig = pcm[u * UTT_SAMPLES:(u + 1) * UTT_SAMPLES]
        post = oracle.process_pitch(oracle.pitch(sig))
        feats = base[fofs[u]:fofs[u + 1]]
        if index == 4:
            energy = oracle.features('energy', sig, dither=dither)
            weights = oracle.vad(energy.astype(np.float32).reshape(-1, 1))
            return feats, post, oracle.cmvn_accumulate(
                feats, weights.reshape(-1).astype(np.float32))
        return feats, post, None
    with concurrent.futures.ThreadPoolExecutor(cores) as pool:
        parts = list(pool.map(one, range(n)))
    if index == 4:
        for s0 in range(0, n, UTTS_PER_SPEAKER):
            group = parts[s0:s0 + UTTS_PER_SPEAKER]
            stats = sum(p[2] for p in group)
            for feats, post, _ in group:
                np.hstack((oracle.deltas(oracle.cmvn_apply(feats, stats)),
                           post[:feats.shape[0]]))
    return int(fofs[-1])


def cpu_baseline(index, pcm_host, nutts_avail, dither, budget_s=12.0):
    """Times the CPU oracle port on a bounded sample of the same workload:
    about `budget_s` seconds of passes over the first n utterances"""
    cores = os.cpu_count() or 1
    probe = min(nutts_avail, max(cores, 8))
    dt, frames = oracle_pass(index, pcm_host, probe, dither, cores)
    rate = frames / dt
    n = int(min(nutts_avail, max(probe, rate * budget_s / FRAMES_PER_UTT)))
    reps = int(max(1, min(64, round(budget_s * rate / (n * FRAMES_PER_UTT)))))
    times, frames = [], 0
    for _ in range(reps):
        d, f = oracle_pass(index, pcm_host, n, dither, cores)
        times.append(d)
        frames += f
    total = float(sum(times))
    return {'value': frames / total, 'unit': 'frames/s', 'cores': cores,
            'kind': 'port',
            'sample': f'{reps} pass(es) over the first {n} utterances of the '
                      f'synthetic corpus ({frames} frames) in {total:.2f} s, '
                      f'C oracle, OpenMP / threads over utterances'}, n, times


def reference_arm(args, config):
    """--impl reference: the reference's own CPU implementation cannot be
    imported (pykaldi is absent): the arm times the
    oracle port on all host cores.  Nothing of libsnb is loaded."""
    import oracle
    oracle.build()
    nutts = min(args.utts or 2048, 2048)
    pcm = synth_host(0, nutts)
    cores = os.cpu_count() or 1
    # one step = one pass over n utterances, n sized for about 8 s
    probe = min(nutts, max(cores, 8))
    dt, frames = oracle_pass(args.config, pcm, probe, args.dither, cores)
    n = int(min(nutts, max(probe, frames / dt * 8.0 / FRAMES_PER_UTT)))
    times, frames = [], 0
    for step in range(args.warmup + args.steps):
        d, f = oracle_pass(args.config, pcm, n, args.dither, cores)
        if step >= args.warmup:
            times.append(d)
            frames += f
    value = frames / float(sum(times))
    config = dict(config, cpu_sample_utterances=n,
                  sharding='rank 0 only, all host cores')
    base = {'value': value, 'unit': 'frames/s', 'cores': cores,
            'kind': 'port',
            'sample': f'each step = one pass over the first {n} utterances '
                      f'of the synthetic corpus ({n * FRAMES_PER_UTT} '
                      f'frames), C oracle, OpenMP / threads over utterances'}
    return {
        'impl': 'reference', 'metric': config['metric'], 'value': value,
        'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': float(np.mean(times) * 1e3),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': config,
        'cpu_baseline': base,
        'e2e': {'value': value, 'unit': 'frames/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}


# --------------------------------------------------------------------------
# host placement
# --------------------------------------------------------------------------
def bind_to_gpu_node(local_rank, world, torch):
    """Runs this rank (and first-touches its pinned buffers) on the CPUs of
    its GPU's NUMA node, split between the ranks that share the node"""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        domain = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist' % (
            domain, bus, dev)
        cpus = []
        for part in open(path).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return None
        sharing = [r for r in range(world) if _cpulist(r, torch) == cpus]
        k, m = sharing.index(local_rank), len(sharing)
        mine = allowed[k * len(allowed) // m:(k + 1) * len(allowed) // m]
        if mine:
            os.sched_setaffinity(0, mine)
        return {'cpus': len(mine), 'node_cpus': len(allowed),
                'ranks_on_node': m}
    except Exception:
        return None


def _cpulist(rank, torch):
    try:
        p = torch.cuda.get_device_properties(rank)
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist' % (
            p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = []
        for part in open(path).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus += list(range(int(lo), int(hi or lo) + 1))
        return cpus
    except Exception:
        return None


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        if rank != 0:
            return
        _, info = config_info(args)
        print(json.dumps(reference_arm(args, info)))
        return

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG', 'WARN')
        dist.init_process_group(
            'nccl', device_id=torch.device('cuda', local_rank))
    placement = bind_to_gpu_node(local_rank, world, torch)
    from shennong_b200 import _lib, engine, stream

    pipe, info = make_config(args.config, args.dither)
    nutts = args.utts or info['utts']
    config = describe(args, info, nutts, world)
    plans = pipe._plans()
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- corpus: device-generated, the first utterances host-generated -----
    # (those are the ones the CPU arms and the parity check read)
    nhost = min(nutts, 2048 if (world == 1 and not args.no_cpu) else 8)
    host_head = synth_host(rank * nutts, nhost)
    pcm_dev = torch.cat([synth_pcm_device(nutts, rank, torch),
                         torch.zeros(64, dtype=torch.int16, device='cuda')])
    pcm_dev[:nhost * UTT_SAMPLES].copy_(torch.from_numpy(host_head))
    starts = np.arange(nutts, dtype=np.int64) * UTT_SAMPLES
    lengths = np.full(nutts, UTT_SAMPLES, dtype=np.int64)
    speakers = ['spk%05d' % (u // UTTS_PER_SPEAKER) for u in range(nutts)]

    # ---- chunks of the device-resident step (one chunk when N = 1) ---------
    nchunks = 1
    if (world > 1 and args.gather != 'none') or args.force_chunks:
        nchunks = max(1, min(args.gather_chunks, nutts // UTTS_PER_SPEAKER))
        if pipe.pitch is not None:
            # the pitch tracker follows one utterance per warp: a chunk must
            # fill a round of it (4 736 utterances) or the kernel runs empty
            wave = int(L.snb_pitch_wave_utts(plans['pitch'].handle))
            nchunks = max(1, min(nchunks, -(-nutts // wave)))
    per = -(-nutts // nchunks)
    per = -(-per // UTTS_PER_SPEAKER) * UTTS_PER_SPEAKER     # whole speakers
    bounds = [(b, min(b + per, nutts)) for b in range(0, nutts, per)]
    chunks = []
    total_frames = 0
    out = None
    for b, e in bounds:
        packed = engine.PackedAudio.from_packed(
            None, starts[b:e] - starts[b], lengths[b:e],
            dev=pcm_dev[int(starts[b]):])
        batches = pipe.make_batches(plans, packed)
        rows = batches['feat'].total_frames
        chunks.append(dict(packed=packed, batches=batches, rows=rows,
                           row0=total_frames, spk=speakers[b:e]))
        total_frames += rows
    base = (None if pipe.simple else torch.empty(
        (total_frames, pipe.base_dim), dtype=torch.float32, device='cuda'))
    modes = ('ce', 'bulk', 'stores', 'nccl')
    collecting = (world > 1 and args.gather != 'none') or (
        args.force_chunks and args.gather in modes)
    hybrid_ok = (pipe.delta is not None and pipe.pitch is None
                 and pipe.cmvn != 'speaker')
    links = (world - 1) * args.gather_fanout
    C = {'coll': None, 'how': None, 'base_chunks': 0, 'ctas': 0}

    def default_base_chunks():
        # with 8 ranks the all-gather of the final rows outlasts the
        # extraction: 5 of 8 chunks travel as base rows
        if not hybrid_ok:
            return 0
        if args.gather_base_chunks >= 0:
            return args.gather_base_chunks
        return (5 * len(chunks) + 4) // 8 if links >= 6 else 0

    def set_collection(how, nb):
        from shennong_b200.distributed import ChunkCollector
        if C['coll'] is not None:
            C['coll'].close()
            C['coll'] = None
            torch.cuda.empty_cache()
        ctas = args.gather_ctas
        if how == 'stores' and not ctas:
            ctas = min(296, 74 * max(world - 1, 1))
        C.update(how=how, base_chunks=nb, ctas=ctas)
        C['coll'] = ChunkCollector(
            pipe, [c['batches']['feat'].frame_offsets for c in chunks],
            how=how, base_chunks=nb, ctas=ctas,
            fanout=args.gather_fanout if how != 'nccl' else 1)
        # the rows are produced IN the gather buffer (own block of every
        # chunk): the collection writes the peers only
        C['outs'] = [C['coll'].out_view(k) for k in range(len(chunks))]

    out = None
    if collecting:
        set_collection('ce' if args.gather == 'auto' else args.gather,
                       default_base_chunks())
    else:
        out = torch.empty((total_frames, pipe.out_dim), dtype=torch.float32,
                          device='cuda')
        C['outs'] = [out[c['row0']:c['row0'] + c['rows']] for c in chunks]

    ev_feat = []

    def step(timed, seed):
        coll, outs = C['coll'], C['outs']
        for k, c in enumerate(chunks):
            r0, r1 = c['row0'], c['row0'] + c['rows']
            if timed:
                pair = (torch.cuda.Event(enable_timing=True),
                        torch.cuda.Event(enable_timing=True))
                ev_feat.append(pair)
                engine.feature_events = pair
            bbuf = coll.base_view(k) if coll is not None else None
            if bbuf is None and base is not None:
                bbuf = base[r0:r1]
            pipe.run_device(
                c['packed'], speakers=c['spk'] if pipe.cmvn == 'speaker'
                else None, seed=seed, out=outs[k], plans=plans,
                base_buf=bbuf, batches=c['batches'],
                norm_out=coll.norm_view(k) if coll is not None else None)
            engine.feature_events = None
            if coll is not None:
                coll.collect(k)
        if coll is not None:
            coll.finish()                # the step ends with the collection

    def trial_ms(nsteps=3):
        """max over ranks of the mean step time (one untimed step first)"""
        step(False, 7)
        barrier()
        t0, t1 = (torch.cuda.Event(enable_timing=True),
                  torch.cuda.Event(enable_timing=True))
        t0.record()
        for i in range(nsteps):
            step(False, 8 + i)
        t1.record()
        barrier()
        t = torch.tensor([t0.elapsed_time(t1) / nsteps], device='cuda',
                         dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- --gather auto: the collection method is chosen by measurement ------
    # (before the warm-up; every rank sees the same max-over-ranks times)
    calibration = None
    if collecting and args.gather == 'auto':
        calibration = {}
        nb0 = default_base_chunks()
        for how in ('ce', 'bulk', 'nccl'):
            if how != C['how'] or nb0 != C['base_chunks']:
                set_collection(how, nb0)
            calibration['%s/%d' % (how, nb0)] = trial_ms()
        best = min(calibration, key=calibration.get).split('/')[0]
        if hybrid_ok and args.gather_base_chunks < 0 and links >= 3:
            # and the number of base-row chunks: walk away from the default
            # while the step gets faster
            for direction in (1, -1):
                nb, last = nb0, calibration['%s/%d' % (best, nb0)]
                while 0 <= nb + direction <= len(chunks) - 1:
                    nb += direction
                    set_collection(best, nb)
                    t = calibration['%s/%d' % (best, nb)] = trial_ms()
                    if t >= last:
                        break
                    last = t
        how, nb = min(calibration, key=calibration.get).split('/')
        if (how, int(nb)) != (C['how'], C['base_chunks']):
            set_collection(how, int(nb))

    for w in range(args.warmup):
        step(False, 100 + w)
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
        step(True, 1000 + i)
    e1.record()
    barrier()
    t1 = time.time()
    launches = L.snb_launch_count() - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    ms_feat = float(np.sum([a.elapsed_time(b) for a, b in ev_feat])
                    / max(args.steps, 1))
    times = torch.tensor([ms_total, ms_feat], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_feat = float(times[0]), float(times[1])
    ms_per_step = ms_total / args.steps
    value = world * total_frames / (ms_per_step * 1e-3)

    # ---- collection alone (N > 1): what the overlap hides --------------------
    gather = None
    if C['coll'] is not None:
        coll = C['coll']
        nck = len(chunks)
        g0, g1 = (torch.cuda.Event(enable_timing=True),
                  torch.cuda.Event(enable_timing=True))
        barrier()
        g0.record()
        for _ in range(3):
            for k in range(nck):
                coll.collect(k)
            coll.finish()
        g1.record()
        barrier()
        gt = torch.tensor([g0.elapsed_time(g1) / 3], device='cuda',
                          dtype=torch.float64)
        if world > 1:
            dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        out = torch.cat(C['outs'])
        gathered = [coll.result(k) for k in range(nck)]
        nbytes = int(out.numel() * 4)
        # bytes a rank sends to EACH peer per step
        as_base = set(coll.base_ids)
        link = sum(c['rows'] * 4 * (pipe.base_dim if k in as_base
                                    else pipe.out_dim)
                   for k, c in enumerate(chunks))
        fan = args.gather_fanout if C['how'] != 'nccl' else 1
        what = {'ce': 'pushed into the result buffers of the other ranks '
                '(CUDA IPC peer memory over NVLink) by the copy engines '
                '(snb_gather_rows_ce)',
                'bulk': 'pushed into the result buffers of the other ranks '
                '(CUDA IPC peer memory over NVLink) by one-warp CTAs driving '
                'the TMA unit (snb_gather_rows_bulk)',
                'stores': 'pushed into the result buffers of the other ranks '
                '(CUDA IPC peer memory over NVLink) by a kernel of 16-byte '
                'stores (snb_gather_rows)',
                'nccl': 'all-gathered chunk by chunk (NCCL over NVLink)'}
        gather = {'in_step': True, 'chunks': nck,
                  'alone_ms': float(gt[0]), 'bytes_per_rank': nbytes,
                  'link_bytes_per_peer': int(link),
                  'recv_gbs_per_rank':
                  (world - 1) * fan * link / (float(gt[0]) * 1e-3) / 1e9,
                  'how': C['how'], 'fanout': fan, 'ctas': C['ctas'],
                  'chunks_as_base_rows': len(as_base),
                  'calibration_ms_per_step': calibration,
                  'api': 'distributed.ChunkCollector: the rows of a chunk are '
                         + what[C['how']] + ' on a second stream while the '
                         'next chunk is computed'
                         + ('; %d chunks travel as base rows + normalisation '
                            'table (a third of the bytes) and every receiver '
                            'redoes the normalise + delta launch on them '
                            '(bit-identical rows)' % len(as_base)
                            if as_base else '')
                         + '; `alone_ms` is the same collection without the '
                         'extraction; with --gather auto the method and the '
                         'number of base-row chunks are the fastest of the '
                         'trials in `calibration_ms_per_step` (run before '
                         'the warm-up)'}
        # the gathered result of the last step against the local rows
        mine = torch.cat([g[rank] for g in gathered])
        gather['own_rows_intact'] = bool(torch.equal(mine, out))
        # and every other rank's rows: float64 checksums of the blocks
        sums = torch.zeros(world, dtype=torch.float64, device='cuda')
        sums[rank] = out.double().sum()
        if world > 1:
            dist.all_reduce(sums)
        got = torch.stack([torch.cat([g[r] for g in gathered]).double().sum()
                           for r in range(world)])
        okay = torch.tensor([float(torch.equal(got, sums))], device='cuda')
        if world > 1:
            dist.all_reduce(okay, op=dist.ReduceOp.MIN)
        gather['rows_of_all_ranks_match_checksums'] = bool(okay.item() == 1.0)

    # ---- parity check on rows of every rank (dither 0, outside the clock) ---
    check = parity_check(args, pipe, info, pcm_dev, host_head, rank, world,
                         nutts, torch, dist, engine)

    # ---- end-to-end: pinned host PCM in, pinned host features out ----------
    e2e = None
    if not args.no_e2e:
        e2e = end_to_end(args, pipe, pcm_dev, out, starts, lengths, speakers,
                         total_frames, world, barrier, torch, dist, stream)
        if placement is not None:
            e2e['host_placement'] = placement

    # ---- the reference-facing API on WAV files -------------------------------
    api = None
    if not args.no_api:
        api = api_leg(args, info, rank, world, barrier, torch, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    kernel_bpf = 320 + 4 * pipe.base_dim
    feat_gbs = total_frames * kernel_bpf / (ms_feat * 1e-3) / 1e9
    roofline = {
        'bound': 'hbm', 'kernel': 'fused_features_512_kernel',
        'achieved': feat_gbs, 'peak': peak, 'unit': 'GB/s',
        'frac': feat_gbs / peak,
        'traffic': (int(total_frames * NCU_TRAFFIC_BYTES_PER_FRAME)
                    if args.config == 2 else None),
        'traffic_source': TRAFFIC_SOURCE if args.config == 2 else None,
        'peak_source': peak_src,
        'algorithmic_bytes_per_launch': int(
            total_frames * kernel_bpf / len(chunks)),
        'launches_per_step': len(chunks),
        'kernel_ms': ms_feat,
        'kernel_share_of_step': ms_feat / ms_per_step,
        'note': ('the chain is instruction-issue / shared-memory bound (~17 '
                 'kflop/frame, AI ~45 flop/B), not HBM bound: see '
                 'fp32_tflops and profiles/'),
        'fp32_tflops': total_frames * FLOPS_PER_FRAME / (ms_feat * 1e-3) / 1e12,
        'pipeline_gbs': total_frames * (320 + 4 * pipe.out_dim)
        / (ms_per_step * 1e-3) / 1e9,
    }
    cpu = None
    if not args.no_cpu and world == 1:       # (rank 0 at N = 1 only: at N > 1
        # the rank is bound to its share of the host cores)
        cpu, _, _ = cpu_baseline(args.config, host_head, nhost, args.dither)

    result = {
        'metric': info['metric'], 'value': value, 'unit': 'frames/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': config, 'clocks': clocks,
        'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
        'cpu_baseline': cpu, 'check': check,
    }
    if api is not None:
        result['e2e_api'] = api
    if gather is not None:
        result['gather'] = gather
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def config_info(args):
    """description of the configuration without touching the GPU library"""
    names = {1: ('filterbank-40 frames/sec', 1000,
                 'BASELINE configs[1]: FilterbankProcessor 40-mel'),
             2: ('MFCC frames/sec', 10000,
                 'BASELINE configs[2]: MfccProcessor (13 ceps, 23 mel, '
                 'reference defaults) + CMVN per utterance (norm_vars) + '
                 'DeltaPostProcessor(order=2, window=2)'),
             3: ('PLP + Kaldi pitch frames/sec', 6250,
                 'BASELINE configs[3]: PlpProcessor + KaldiPitchProcessor '
                 '(+ post-processing) pipeline, 50 000 utterances over 8 GPUs'),
             4: ('full pipeline frames/sec', 4500,
                 'BASELINE configs[4]: speech-features full pipeline '
                 '(filterbank + Kaldi pitch + delta + CMVN by speaker with '
                 'VAD), 100 h = 36 000 utterances over 8 GPUs')}
    metric, utts, workload = names[args.config]
    info = dict(metric=metric, utts=utts, workload=workload)
    nutts = args.utts or utts
    world = int(os.environ.get('WORLD_SIZE', 1))
    return None, dict(describe(args, info, nutts, world), metric=metric)


def describe(args, info, nutts, world):
    return {
        'workload': info['workload'] + f' on {nutts} synthetic 16 kHz 10 s '
        'utterances per GPU',
        'baseline_config': args.config,
        'utterances_per_gpu': nutts, 'frames_per_utterance': FRAMES_PER_UTT,
        'dither': args.dither,
        'sharding': f'utterances x{world}, no data-path collective; '
        'collection (all-gather of the rows) inside the step for N > 1',
        'l2': 'inputs (320 KB per utterance, GBs per GPU) larger than the '
        '126 MB L2',
    }


def parity_check(args, pipe, info, pcm_dev, host_head, rank, world, nutts,
                 torch, dist, engine):
    """dither-0 rows of the first utterances of EVERY rank (gathered over
    NCCL for N > 1) against the CPU oracle, on rank 0"""
    import oracle
    ncheck = min(8, nutts)
    cpipe, _ = make_config(args.config, 0.0)
    if cpipe.pitch is not None:
        cpipe.pitch[1].delta_pitch_noise_stddev = 0
    if cpipe.energy is not None:
        cpipe.energy.dither = 0
    packed = engine.PackedAudio.from_packed(
        None, np.arange(ncheck, dtype=np.int64) * UTT_SAMPLES,
        np.full(ncheck, UTT_SAMPLES, dtype=np.int64), dev=pcm_dev)
    rows, _, _, _ = cpipe.run_device(
        packed, speakers=['s'] * ncheck if cpipe.cmvn == 'speaker' else None)
    if world > 1:
        full = torch.empty((world * rows.shape[0], rows.shape[1]),
                           dtype=torch.float32, device='cuda')
        dist.all_gather_into_tensor(full, rows.contiguous())
    else:
        full = rows
    if rank != 0:
        return None
    full = full.cpu().numpy().reshape(world, ncheck, FRAMES_PER_UTT, -1)
    kind, kw = info['oracle']
    worst, worst_pitch = 0.0, 0.0
    for r in range(world):
        head = host_head if r == 0 else synth_host(r * nutts, ncheck)
        sigs = [head[u * UTT_SAMPLES:(u + 1) * UTT_SAMPLES]
                for u in range(ncheck)]
        feats = [oracle.features(kind, sig, dither=0, **kw) for sig in sigs]
        # per-column bound: 1e-4 of the largest base coefficient; a CMVN with
        # variance normalisation divides column d by its standard deviation
        scale = max(float(np.abs(f).max()) for f in feats)
        if args.config == 2:
            want = [oracle.deltas(oracle.cmvn_apply(
                f, oracle.cmvn_accumulate(f))) for f in feats]
            bound = [np.tile(1e-4 * scale / f.std(axis=0), 3) for f in feats]
        elif args.config == 4:
            stats = 0
            for sig, f in zip(sigs, feats):
                e = oracle.features('energy', sig, dither=0)
                w = oracle.vad(e.astype(np.float32).reshape(-1, 1))
                stats = stats + oracle.cmvn_accumulate(
                    f, w.reshape(-1).astype(np.float32))
            want = [oracle.deltas(oracle.cmvn_apply(f, stats)) for f in feats]
            std = np.concatenate(feats).std(axis=0)
            bound = [np.tile(1e-4 * scale / std, 3)] * ncheck
        else:
            want = feats
            bound = [np.full(f.shape[1], 1e-4 * scale) for f in feats]
        for u in range(ncheck):
            got = full[r, u]
            d = want[u].shape[1]
            ratio = (np.abs(got[:, :d] - want[u]) / bound[u][None, :]).max()
            worst = max(worst, float(ratio))
            if pipe.pitch is not None:
                post = oracle.process_pitch(oracle.pitch(sigs[u]))
                worst_pitch = max(worst_pitch, float(
                    np.abs(got[:, d:] - post).max()))
    return {'ranks': world, 'utterances_per_rank': ncheck,
            'max_err_over_bound': worst,
            'bound': '1e-4 * max|base coefficient| per column, divided by '
            'the column\'s standard deviation after a variance-normalising '
            'CMVN',
            'max_abs_err_pitch_columns':
            worst_pitch if pipe.pitch is not None else None,
            'ok': bool(worst <= 1.0 and worst_pitch <= 1e-4),
            'note': 'dither 0, oracle = CPU restatement (oracle/); pitch '
            'columns come from the bit-exact state sequence'}


def end_to_end(args, pipe, pcm_dev, out, starts, lengths, speakers,
               total_frames, world, barrier, torch, dist, stream):
    nutts = len(lengths)
    host_pcm = torch.empty(pcm_dev.numel(), dtype=torch.int16, pin_memory=True)
    host_pcm.copy_(pcm_dev)
    out_host = torch.empty((total_frames, pipe.out_dim), dtype=torch.float32,
                           pin_memory=True)
    torch.cuda.synchronize()
    # ceilings of this box: plain pinned copies, one direction each, then both
    # directions at once -- every rank at the same time (what the host can
    # deliver to N GPUs together)
    c0, c1 = (torch.cuda.Event(enable_timing=True),
              torch.cuda.Event(enable_timing=True))
    scratch = torch.empty_like(pcm_dev)
    barrier()
    c0.record(); scratch.copy_(host_pcm, non_blocking=True); c1.record()
    barrier()
    h2d_gbs = host_pcm.numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    c0.record(); out_host.copy_(out, non_blocking=True); c1.record()
    barrier()
    d2h_gbs = out.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    s2 = torch.cuda.Stream()
    barrier()
    t_b = time.perf_counter()
    scratch.copy_(host_pcm, non_blocking=True)
    with torch.cuda.stream(s2):
        out_host.copy_(out, non_blocking=True)
    barrier()
    both_s = time.perf_counter() - t_b
    tb = torch.tensor([both_s], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    both_s = float(tb[0])
    del scratch
    spk = speakers if pipe.cmvn == 'speaker' else None
    nrep = max(2, min(args.steps, 5))
    for _ in range(3):                                           # warm-up
        pipe.run_host(host_pcm, starts, lengths, out_host=out_host,
                      chunk_utts=args.chunk_utts or None, speakers=spk)
    barrier()
    rep_ms = []
    t_e0 = time.perf_counter()
    for _ in range(nrep):
        t_r = time.perf_counter()
        pipe.run_host(host_pcm, starts, lengths, out_host=out_host,
                      chunk_utts=args.chunk_utts or None, speakers=spk)
        rep_ms.append((time.perf_counter() - t_r) * 1e3)
    barrier()
    dt = (time.perf_counter() - t_e0) / nrep
    tt = torch.tensor([dt], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])
    return {'value': world * total_frames / dt, 'unit': 'frames/s',
            'h2d_bytes_per_step': int(nutts * UTT_SAMPLES * 2),
            'd2h_bytes_per_step': int(total_frames * pipe.out_dim * 4),
            'ms_per_step': dt * 1e3,
            'ms_each_step': [round(t, 2) for t in rep_ms],
            'api': 'FusedPipeline.run_host (stream.StreamRunner: chunked '
                   f'H2D / compute / D2H, {pipe._runner.chunk_utts} '
                   'utterances per chunk; every rank returns the rows of its '
                   'shard)',
            'pcie_h2d_gbs': h2d_gbs, 'pcie_d2h_gbs': d2h_gbs,
            'host_ceiling_ms': both_s * 1e3,
            'fraction_of_host_ceiling': both_s / dt,
            'ceiling': 'the same H2D + D2H bytes as plain concurrent pinned '
                       'copies, all ranks at once, no compute'}


def api_leg(args, info, rank, world, barrier, torch, dist):
    """The calls a user of the reference makes, on WAV files: process_all of
    the main features processor and pipeline.extract_features of the
    configuration (sharded over the ranks; every rank gets every row)"""
    import scipy.io.wavfile
    from shennong_b200 import Utterances, pipeline
    from shennong_b200.processor import MfccProcessor
    n = args.api_utts
    root = os.path.join(tempfile.gettempdir(), 'snb_bench_wavs_%d' % n)
    if rank == 0:
        os.makedirs(root, exist_ok=True)
        pcm = synth_host(0, n)
        for u in range(n):
            path = os.path.join(root, 'u%05d.wav' % u)
            if not os.path.exists(path):
                scipy.io.wavfile.write(
                    path, SAMPLE_RATE,
                    pcm[u * UTT_SAMPLES:(u + 1) * UTT_SAMPLES])
    barrier()
    utts = Utterances([('u%05d' % u, os.path.join(root, 'u%05d.wav' % u),
                        'spk%03d' % (u // UTTS_PER_SPEAKER))
                       for u in range(n)])
    cores = len(os.sched_getaffinity(0))
    results = {}
    proc = MfccProcessor(dither=args.dither)
    config = pipeline.get_default_config(
        'mfcc', with_cmvn=True, with_delta=True)
    config['mfcc']['dither'] = args.dither

    def timed(name, call, dim):
        call()                                                  # warm-up
        barrier()
        t0 = time.perf_counter()
        feats = call()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        frames = sum(f.nframes for f in feats.values())
        results[name] = {'value': frames / float(tt[0]), 'unit': 'frames/s',
                         'ms': float(tt[0]) * 1e3, 'utterances': len(feats),
                         'frames': frames, 'dim': dim, 'njobs': cores}
    timed('MfccProcessor.process_all',
          lambda: proc.process_all(utts, njobs=cores), 13)
    timed('pipeline.extract_features (mfcc + cmvn by speaker with vad + delta)',
          lambda: pipeline.extract_features(config, utts, njobs=cores), 39)
    # floor of anything that starts from files: the bytes of this rank's
    # share read from the page cache into pinned memory, nothing else
    from shennong_b200 import stream
    from shennong_b200.distributed import shard_utterances
    ulist = [utts[k] for k in utts.by_name().keys()]
    items, lengths, _ = stream.audio_items(ulist)
    mine = shard_utterances(lengths, world)[rank]
    source = stream.AudioSource([items[i] for i in mine], lengths[mine],
                                workers=cores)
    step = 512
    staging = torch.empty(source.span(0, min(step, len(mine))),
                          dtype=torch.int16, pin_memory=True)
    for rep in range(2):
        barrier()
        t0 = time.perf_counter()
        for b in range(0, len(mine), step):
            source.window(b, min(b + step, len(mine)), staging)
        barrier()
        dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    results['read_floor'] = {
        'ms': float(tt[0]) * 1e3,
        'gbs_per_rank': float(lengths[mine].sum()) * 2 / float(tt[0]) / 1e9,
        'what': 'readinto of the same WAV payloads into pinned staging by '
                f'{cores} threads, no GPU work'}
    results['note'] = (
        f'{n} WAV files of 10 s read from {tempfile.gettempdir()} inside the '
        'timed call (mono 16-bit PCM payloads go straight into pinned '
        'staging), rows of all ranks gathered to every rank, one Features '
        'per utterance')
    if rank == 0 and os.environ.get('SNB_BENCH_KEEP_WAVS') is None:
        shutil.rmtree(root, ignore_errors=True)
    return results


if __name__ == '__main__':
    main()
