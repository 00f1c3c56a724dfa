"""Every public module imports in a fresh interpreter, in any order"""

import subprocess
import sys

import pytest

from conftest import ROOT

MODULES = [
    'shennong_b200', 'shennong_b200.processor', 'shennong_b200.postprocessor',
    'shennong_b200.fused', 'shennong_b200.processor.pitch_kaldi',
    'shennong_b200.postprocessor.cmvn', 'shennong_b200.frames',
    'shennong_b200.window', 'shennong_b200.engine']


@pytest.mark.parametrize('module', MODULES)
def test_import_alone(module):
    subprocess.run(
        [sys.executable, '-c', f'import {module}'], check=True, cwd=ROOT)
