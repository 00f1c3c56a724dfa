"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every
kernel of the bench step plus PLP, pitch and a VTLN batch, with more tiles than
resident CTAs so that the descriptor ring and both PCM buffers of the fused
kernel wrap around.

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    from conftest import synth_utterance
    from shennong_b200 import engine
    from shennong_b200.fused import FusedPipeline
    from shennong_b200.postprocessor import DeltaPostProcessor
    from shennong_b200.processor import (
        KaldiPitchPostProcessor, KaldiPitchProcessor, MfccProcessor, PlpProcessor)
    lengths = [160000] * 9 + [401, 9000, 31999, 123457]
    sigs = [synth_utterance(i, n) for i, n in enumerate(lengths)]
    packed = engine.PackedAudio(sigs)
    for pipe, kw in (
            (FusedPipeline(MfccProcessor(), delta=DeltaPostProcessor(), cmvn='utterance'), {}),
            (FusedPipeline(MfccProcessor(dither=0, snip_edges=False)), {}),
            (FusedPipeline(PlpProcessor(rasta=True)), {}),
            (FusedPipeline(MfccProcessor(dither=0)),
             {'warps': np.linspace(0.9, 1.1, len(sigs)).astype(np.float32)}),
            (FusedPipeline(MfccProcessor(), pitch=(KaldiPitchProcessor(), KaldiPitchPostProcessor())), {})):
        out, offs, _, _ = pipe.run_device(packed, **kw)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        print(type(pipe.processor).__name__, tuple(out.shape), 'ok', flush=True)


if __name__ == '__main__':
    main()
