"""Base class of the features post-processors

    Features --> FeaturesPostProcessor --> Features

(counterpart of shennong/postprocessor/base.py)
"""

import abc
import copy

from shennong_b200.processor.base import FeaturesProcessor


class FeaturesPostProcessor(FeaturesProcessor):
    """Base class of all features post-processors"""
    @abc.abstractmethod
    def process(self, features):
        """Returns features post-processed from input `features`"""

    def get_properties(self, features):
        properties = copy.deepcopy(features.properties)
        properties[self.name] = self.get_params()
        properties.setdefault('pipeline', []).append(
            {'name': self.name, 'columns': [0, self.ndims - 1]})
        return properties
