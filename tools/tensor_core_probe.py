"""Tensor-core probe for the DFT stage

north_star allows tensor cores "only if ... ncu shows tensor-pipe utilisation
wins over the warp path".  This measures the upper bound of the dense-DFT
formulation on the bench corpus, with the library GEMM as the tensor-core
engine (cuBLAS bf16 / fp16 through torch.matmul: the best case a hand-written
tcgen05 kernel could approach, measured instead of estimated):

    frames [F, 400] (int16 samples, exact split in low-precision terms)
      x  M [400, 514]  (DC removal, pre-emphasis, Povey window and the 512-point
                        real DFT folded into one matrix: all linear)
    -> power spectrum -> 23 mel energies -> log          (fp32, torch)

and compares (a) frames/s of the GEMM passes ALONE (no framing, no epilogue,
no dither -- everything else of the fused kernel is free in this bound) with
the fused CUDA-core kernel doing the WHOLE chain, (b) the log-mel error against
a float64 evaluation, for 1-, 2- and 3-term splits.

    python tools/tensor_core_probe.py [--utts 2000]
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dft_matrix():
    """[400, 514] float64: x (raw frame) -> interleaved (Re, Im) of the
    512-point DFT of window * preemph(x - mean(x))"""
    W, N = 400, 512
    n = np.arange(W)
    a = 2 * np.pi / (W - 1)
    window = (0.5 - 0.5 * np.cos(a * n)) ** 0.85
    # y = P (I - 11^T / W) x ; z = diag(window) y
    dc = np.eye(W) - np.ones((W, W)) / W
    pre = np.eye(W)
    pre[n[1:], n[:-1]] -= 0.97
    pre[0, 0] -= 0.97
    lin = window[:, None] * (pre @ dc)                  # [W, W]
    k = np.arange(N // 2 + 1)
    ang = -2 * np.pi * np.outer(n, k) / N
    F = np.empty((W, 2 * len(k)))
    F[:, 0::2] = np.cos(ang)
    F[:, 1::2] = np.sin(ang)
    return lin.T @ F                                     # x^T M


def split(x, dtype, terms, torch):
    parts, rest = [], x.double()
    for _ in range(terms):
        p = rest.to(dtype)
        parts.append(p)
        rest = rest - p.double()
    return parts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--utts', type=int, default=2000)
    args = ap.parse_args()
    import torch
    import bench
    from shennong_b200 import engine
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.processor import FilterbankProcessor

    n = args.utts
    pcm = torch.cat([bench.synth_pcm_device(n, 0, torch),
                     torch.zeros(64, dtype=torch.int16, device='cuda')])
    frames_per = bench.FRAMES_PER_UTT
    # framing by strided view: [n, 998, 400] int16 -> float
    x = pcm[:n * bench.UTT_SAMPLES].view(n, bench.UTT_SAMPLES)
    frames = x.unfold(1, 400, 160).reshape(-1, 400)      # view + copy below
    F = frames.shape[0]
    M64 = torch.from_numpy(dft_matrix()).cuda()

    results = {}
    for name, dtype in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
        for xt, mt in ((1, 1), (2, 2), (3, 2), (2, 3)):
            # exact split of the int16 samples / rounded split of the matrix
            xs = split(frames.float(), dtype, xt, torch)
            ms = split(M64, dtype, mt, torch)
            pairs = [(i, j) for i in range(xt) for j in range(mt)
                     if i + j < max(xt, mt)]
            out = torch.zeros((F, 514), dtype=torch.float32, device='cuda')

            def run():
                out.zero_()
                for i, j in pairs:
                    out.add_(torch.matmul(xs[i], ms[j]).float())
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            # time the GEMM passes alone
            e0, e1 = (torch.cuda.Event(enable_timing=True),
                      torch.cuda.Event(enable_timing=True))
            e0.record()
            for _ in range(3):
                for i, j in pairs:
                    torch.matmul(xs[i], ms[j])
            e1.record()
            torch.cuda.synchronize()
            ms_gemm = e0.elapsed_time(e1) / 3
            # accuracy on a sample of frames: power spectrum against float64
            sel = slice(0, min(F, 200000))
            ref = frames[sel].double() @ M64
            pw_ref = ref[:, 0::2] ** 2 + ref[:, 1::2] ** 2
            got = out[sel].double()
            pw = got[:, 0::2] ** 2 + got[:, 1::2] ** 2
            # 23 mel bins ~ sums of ~10 power bins: use 8-bin boxcar sums as a
            # proxy of the mel energies, log domain, scale-relative error
            def boxes(p):
                return torch.log(torch.clamp(
                    p[:, 1:257].reshape(p.shape[0], 32, 8).sum(-1), min=1e-7))
            lr, lg = boxes(pw_ref), boxes(pw)
            err = float((lr - lg).abs().max() / lr.abs().max())
            results[f'{name} x{xt}-term input, {mt}-term matrix'] = {
                'gemm_passes': len(pairs), 'ms_gemm_only': ms_gemm,
                'frames_per_s_gemm_only': F / (ms_gemm * 1e-3),
                'tflops': len(pairs) * 2 * 400 * 514 * F / (ms_gemm * 1e-3) / 1e12,
                'max_scale_rel_err_log_band_energy': err}
            print(f'{name} input {xt} term(s) matrix {mt}: {len(pairs)} GEMM '
                  f'passes {ms_gemm:8.3f} ms -> {F / (ms_gemm * 1e-3):.3e} '
                  f'frames/s (GEMM only), log-energy err {err:.2e}', flush=True)
    # the fused CUDA-core kernel on the same frames: the WHOLE chain
    pipe = FusedPipeline(FilterbankProcessor(dither=0))
    starts = np.arange(n, dtype=np.int64) * bench.UTT_SAMPLES
    lengths = np.full(n, bench.UTT_SAMPLES, dtype=np.int64)
    packed = engine.PackedAudio.from_packed(None, starts, lengths, dev=pcm)
    plans = pipe._plans()
    for _ in range(3):
        pipe.run_device(packed, plans=plans)
    torch.cuda.synchronize()
    e0, e1 = (torch.cuda.Event(enable_timing=True),
              torch.cuda.Event(enable_timing=True))
    e0.record()
    for _ in range(5):
        pipe.run_device(packed, plans=plans)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    results['fused CUDA-core kernel (whole chain: framing .. log-mel)'] = {
        'ms': ms, 'frames_per_s': F / (ms * 1e-3)}
    print(f'fused kernel, whole chain: {ms:.3f} ms -> {F / (ms * 1e-3):.3e} '
          f'frames/s')
    print(json.dumps(results))


if __name__ == '__main__':
    main()
