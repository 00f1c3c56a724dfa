"""Features post-processors (same names as shennong.postprocessor)"""

from shennong_b200.postprocessor.cmvn import (
    CmvnPostProcessor, SlidingWindowCmvnPostProcessor, apply_cmvn)
from shennong_b200.postprocessor.delta import DeltaPostProcessor
from shennong_b200.postprocessor.vad import VadPostProcessor

__all__ = ['CmvnPostProcessor', 'SlidingWindowCmvnPostProcessor',
           'apply_cmvn', 'DeltaPostProcessor', 'VadPostProcessor']
