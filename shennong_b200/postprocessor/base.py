"""Base class of the features post-processors

    Features --> FeaturesPostProcessor --> Features

(counterpart of shennong/postprocessor/base.py; the class body lives in
shennong_b200.processor.base to keep the two packages free of import cycles)
"""

from shennong_b200.processor.base import FeaturesPostProcessor

__all__ = ['FeaturesPostProcessor']
