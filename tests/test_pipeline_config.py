"""Pipeline configuration surface (CPU): defaults, YAML, validation --
counterpart of the reference's test/test_pipeline.py:40-146"""

import numpy as np
import pytest
import yaml

from shennong_b200 import pipeline
from shennong_b200.pipeline_manager import PipelineManager


def test_valid_features_and_processors():
    assert pipeline.valid_features() == [
        'spectrogram', 'filterbank', 'mfcc', 'plp']
    for name in PipelineManager.valid_processors:
        cls = PipelineManager.get_processor_class(name)
        assert cls.__name__.endswith('Processor')
    for name in ('bottleneck', 'crepe_pitch', 'vtln', 'ubm', 'nope'):
        with pytest.raises(ValueError):
            PipelineManager.get_processor_class(name)


@pytest.mark.parametrize('features', ['spectrogram', 'filterbank', 'mfcc',
                                      'plp'])
def test_default_config(features):
    config = pipeline.get_default_config(features)
    assert list(config.keys()) == [features]
    assert 'sample_rate' not in config[features]
    assert 'htk_compat' not in config[features]
    assert config[features]['dither'] == 1.0
    full = pipeline.get_default_config(
        features, with_pitch='kaldi', with_cmvn=True, with_delta=True)
    assert list(full.keys()) == [features, 'pitch', 'cmvn', 'delta']
    assert full['pitch']['processor'] == 'kaldi'
    assert 'frame_shift' not in full['pitch']
    assert full['pitch']['postprocessing']['add_pov_feature'] is True
    assert full['cmvn']['by_speaker'] and full['cmvn']['with_vad']
    assert full['cmvn']['vad']['energy_threshold'] == 5.0
    assert full['delta'] == {'order': 2, 'window': 2}


def test_default_config_errors():
    with pytest.raises(ValueError):
        pipeline.get_default_config('bottleneck')
    with pytest.raises(ValueError):
        pipeline.get_default_config('mfcc', with_pitch='foo')
    with pytest.raises(ValueError):
        pipeline.get_default_config('mfcc', with_pitch='crepe')
    with pytest.raises(ValueError):
        pipeline.get_default_config('mfcc', with_vtln='simple')


@pytest.mark.parametrize('commented', [True, False])
def test_yaml_round_trip(commented):
    kw = dict(with_pitch='kaldi', with_cmvn=True, with_delta=True)
    text = pipeline.get_default_config(
        'mfcc', to_yaml=True, yaml_commented=commented, **kw)
    assert ('#' in text) == commented
    parsed = yaml.load(text, Loader=yaml.FullLoader)
    config = pipeline.get_default_config('mfcc', **kw)
    assert parsed.keys() == config.keys()
    for key in config:
        for param, value in config[key].items():
            if isinstance(value, dict):
                for p2, v2 in value.items():
                    assert parsed[key][param][p2] == pytest.approx(v2)
            else:
                assert parsed[key][param] == pytest.approx(value)
    if commented:
        assert '# Frame shift in seconds. Default is' in text
        assert '# If false, do normalization by utterance' in text
        assert 'Computing pitch using kaldi' in text
    # the YAML text itself is a valid configuration
    pipeline._init_config(text)


def test_init_config_validation():
    good = pipeline.get_default_config('mfcc', with_cmvn=True)
    assert pipeline._init_config(good)['cmvn']['with_vad'] is True
    with pytest.raises(ValueError, match='invalid keys'):
        pipeline._init_config({'mfcc': {}, 'foo': {}})
    with pytest.raises(ValueError, match='does not define any features'):
        pipeline._init_config({'delta': {}})
    with pytest.raises(ValueError, match='more than one'):
        pipeline._init_config({'mfcc': {}, 'plp': {}})
    with pytest.raises(ValueError, match='not available'):
        pipeline._init_config({'mfcc': {}, 'vtln': {}})
    with pytest.raises(ValueError):
        pipeline._init_config('mfcc: [unclosed')
    # missing options are completed like the reference does
    config = pipeline._init_config({'mfcc': {}, 'cmvn': {}, 'pitch': {}})
    assert config['cmvn'] == {
        'by_speaker': False, 'with_vad': True,
        'vad': PipelineManager.get_processor_params('vad')}
    assert config['pitch'] == {'processor': 'kaldi', 'postprocessing': {}}
    # the caller's dict is not modified
    source = {'mfcc': {}, 'cmvn': {}}
    pipeline._init_config(source)
    assert source == {'mfcc': {}, 'cmvn': {}}


def test_docstrings_feed_yaml_comments():
    doc = PipelineManager.get_docstring('mfcc', 'num_ceps', 13)
    assert doc.startswith('Number of cepstra in MFCC computation')
    assert doc.endswith('Default is 13.')
    assert np.float32(1.0) == pipeline.get_default_config('plp')['plp'][
        'dither']
