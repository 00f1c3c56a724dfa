"""Speech features extraction pipeline (counterpart of shennong/pipeline.py)

    features (+ VTLN warps) -> CMVN (by speaker or utterance, optional VAD)
    -> delta   ‖   pitch -> pitch post-processing      => column concatenation

Same configuration surface (dict / YAML, :func:`get_default_config`) and same
entry point :func:`extract_features`, but the schedule is batched: the
utterances are loaded on host threads (`njobs`), packed, and the whole
two-pass schedule of the reference (pipeline.py:541-567) runs as a short
sequence of fused launches on the device
(:class:`shennong_b200.fused.FusedPipeline`).

>>> from shennong_b200 import pipeline
>>> config = pipeline.get_default_config('mfcc', with_cmvn=True)
>>> sorted(config.keys())
['cmvn', 'mfcc']
"""

import concurrent.futures
import copy
import os
import textwrap

import numpy as np
import yaml

from shennong_b200 import engine, stream
from shennong_b200.features import Features
from shennong_b200.features_collection import FeaturesCollection
from shennong_b200.fused import FusedPipeline
from shennong_b200.logger import get_logger
from shennong_b200.pipeline_manager import PipelineManager
from shennong_b200.utils import get_njobs


def valid_features():
    """The main features available to the pipeline"""
    return PipelineManager.valid_features


def get_default_config(features, to_yaml=False, yaml_commented=True,
                       with_pitch=False, with_cmvn=False, with_delta=False,
                       with_vtln=False):
    """Default configuration of the pipeline for `features`

    Returns a dict, or a YAML string if `to_yaml` (with the parameters
    docstrings as comments if `yaml_commented`).  `with_pitch` is False or
    'kaldi'.  Raises ValueError on invalid arguments; 'crepe' pitch and VTLN
    estimation are not provided by this engine.
    """
    if features not in valid_features():
        raise ValueError('invalid features "{}", must be in {}'.format(
            features, ', '.join(valid_features())))
    if with_pitch not in (False, 'kaldi', 'crepe'):
        raise ValueError(
            f'with_pitch argument must be False, "kaldi" or "crepe" '
            f'but is "{with_pitch}"')
    if with_pitch == 'crepe':
        PipelineManager.get_processor_class('crepe_pitch')   # raises
    if with_vtln not in (False, 'simple', 'full'):
        raise ValueError(
            f'with_vtln argument must be False, "simple" or "full" '
            f'but is "{with_vtln}"')
    if with_vtln:
        PipelineManager.get_processor_class('vtln')          # raises

    params = PipelineManager.get_processor_params
    config = {features: {
        k: v for k, v in params(features).items()
        if k not in ('sample_rate', 'htk_compat')}}
    if with_pitch:
        config['pitch'] = {'processor': with_pitch}
        config['pitch'].update(
            (k, v) for k, v in params('kaldi_pitch').items()
            if k not in ('frame_length', 'frame_shift', 'sample_rate'))
        config['pitch']['postprocessing'] = params('kaldi_pitch_post')
    if with_cmvn:
        config['cmvn'] = {
            'by_speaker': True, 'with_vad': True, 'vad': params('vad')}
    if with_delta:
        config['delta'] = params('delta')
    if to_yaml:
        return _config_to_yaml(config, comments=yaml_commented)
    return config


class _Dumper(yaml.SafeDumper):
    """Keeps the insertion order and knows the numpy scalar types"""


_Dumper.add_representer(
    dict, lambda d, data: d.represent_dict(data.items()))
_Dumper.add_representer(
    np.float32, lambda d, data: d.represent_float(float(data)))
_Dumper.add_representer(
    np.float64, lambda d, data: d.represent_float(float(data)))
_Dumper.add_representer(
    np.bool_, lambda d, data: d.represent_bool(bool(data)))


def _config_to_yaml(config, comments=True):
    """YAML rendering of a configuration, optionally commented with the
    documentation of each parameter (shennong/pipeline.py:315-416)"""
    text = yaml.dump(config, Dumper=_Dumper).strip()
    if not comments:
        return text + '\n'
    out, stack, prev_indent = [], [], 0
    for line in text.split('\n'):
        key = line.split(': ')[0]
        indent = len(key) - len(key.lstrip())
        for _ in range((prev_indent - indent) // 2):
            stack.pop()
        if line.endswith(':'):
            name = line[:-1].strip()
            if name == 'postprocessing':
                name = f'{stack[-1]}_post'
            stack.append(name)
            if name == 'vad' and indent != 4:
                out.append(
                    "  # The vad options are not used if 'with_vad' is false")
            out.append(line)
        else:
            param, default = key.strip(), line.split(': ')[1].strip()
            owner = stack[-1]
            if owner == 'cmvn' and param == 'by_speaker':
                doc = ('If false, do normalization by utterance, '
                       'if true do normalization by speaker.')
            elif owner == 'cmvn' and param == 'with_vad':
                doc = ('If true do normalization only on frames where '
                       'voice activity has been detected, if false do not '
                       'consider voice activity for normalization.')
            elif owner == 'pitch' and param == 'processor':
                doc = f'Computing pitch using {default}'
            elif 'pitch' in owner:
                doc = PipelineManager.get_docstring(
                    'kaldi_' + owner, param, default)
            else:
                doc = PipelineManager.get_docstring(owner, param, default)
            out += [' ' * indent + '# ' + w
                    for w in textwrap.wrap(doc, width=68 - indent)]
            out.append(line)
        prev_indent = indent
    return '\n'.join(out) + '\n'


def _init_config(config, log=get_logger('pipeline', 'warning')):
    """Parses and validates a configuration (dict, YAML string or file)"""
    try:
        if os.path.isfile(config):
            log.debug('loading configuration from %s', config)
            with open(config, 'r') as stream:
                config = stream.read()
    except TypeError:
        pass
    if isinstance(config, str):
        try:
            config = yaml.load(config, Loader=yaml.FullLoader)
        except yaml.YAMLError as err:
            raise ValueError(f'error in configuration: {err}') from None
    config = copy.deepcopy(config)
    known = list(PipelineManager.valid_processors) + [
        'pitch', 'bottleneck', 'vtln', 'ubm']
    unknown = [k for k in config.keys() if k not in known]
    if unknown:
        raise ValueError(
            'invalid keys in configuration: {}'.format(', '.join(unknown)))
    for key in ('bottleneck', 'vtln', 'ubm'):
        if key in config:
            PipelineManager.get_processor_class(key)          # raises
    features = [k for k in config.keys() if k in valid_features()]
    if not features:
        raise ValueError(
            'the configuration does not define any features extraction '
            '(must have one and only one entry of {})'
            .format(', '.join(valid_features())))
    if len(features) > 1:
        raise ValueError(
            'more than one features extraction processors are defined, '
            '(must have one and only one entry of {}): {}'
            .format(', '.join(valid_features()), ', '.join(features)))
    if 'cmvn' in config:
        if 'by_speaker' not in config['cmvn']:
            log.warning(
                'by_speaker option not specified for cmvn, '
                'assuming it is false and doing cmvn by utterance')
            config['cmvn']['by_speaker'] = False
        config['cmvn'].setdefault('with_vad', True)
        if config['cmvn']['with_vad']:
            config['cmvn'].setdefault(
                'vad', PipelineManager.get_processor_params('vad'))
    if 'pitch' in config:
        config['pitch'].setdefault('processor', 'kaldi')
        config['pitch'].setdefault('postprocessing', {})
    steps = []
    if 'pitch' in config:
        steps.append(f'{config["pitch"]["processor"]} pitch')
    if 'delta' in config:
        steps.append('delta')
    if 'cmvn' in config:
        steps.append('cmvn by {}{}'.format(
            'speaker' if config['cmvn']['by_speaker'] else 'utterance',
            ' with vad' if config['cmvn']['with_vad'] else ''))
    log.info('pipeline configured for %s features extraction%s', features[0],
             ' with {}'.format(', '.join(steps)) if steps else '')
    return config


def _init_warps(warps, config, utterances, log):
    features = [k for k in config.keys() if k in valid_features()][0]
    if features == 'spectrogram':
        raise ValueError(f'{features} features do not support VTLN')
    if warps.keys() == utterances.by_name().keys():
        log.info('VTLN warps are defined by utterance')
    elif (not utterances.has_speakers()
          or warps.keys() != utterances.by_speaker().keys()):
        raise ValueError(
            'warps do not match utterances, either by speaker or by utterance')
    else:
        log.info('VTLN warps are defined by speaker')
        warps = {utt.name: warps[utt.speaker] for utt in utterances}
    return {name: float(warp) for name, warp in warps.items()}


class _Carrier:
    """Stands for a Features when only properties and ndims flow through the
    processors' ``get_properties`` (the data itself stays on the device)"""
    def __init__(self, properties, ndims):
        self.properties = properties
        self.ndims = ndims


def extract_features(configuration, utterances, warps=None, njobs=1,
                     log=get_logger('pipeline', 'warning'), gather=True):
    """Extracts the features of all the `utterances`

    Parameters
    ----------
    configuration : dict or str
        Pipeline configuration: a dict, a YAML string or the path to a YAML
        file (see :func:`get_default_config`)
    utterances : Utterances
        The utterances to extract the features on
    warps : dict, optional
        Known VTLN warps indexed by utterance name or by speaker
    njobs : int, optional
        Host threads that read the audio into the pinned staging buffers
        (the extraction itself is batched on the GPU)
    gather : bool, optional
        Only under ``torch.distributed`` (one process per GPU): every rank
        extracts its own shard of the utterances (whole speakers when CMVN
        is by speaker); with `gather` the rows of all ranks are all-gathered
        over NCCL while the extraction proceeds and every rank returns the
        complete collection, else each rank returns its shard.

    The corpus is streamed: utterances are read, uploaded, processed and
    downloaded in chunks, so that neither device memory nor pinned host
    staging grow with the corpus (the reference streams utterance by
    utterance, pipeline.py:541-567); only the result lives in host memory.

    Returns
    -------
    features : FeaturesCollection

    Raises
    ------
    ValueError
        On invalid configuration, utterances or warps
    """
    njobs = get_njobs(njobs, log=log)
    config = _init_config(configuration, log=log)
    log.info('detected format for utterances index is: %s',
             utterances.format(type=str))
    if warps:
        warps = _init_warps(warps, config, utterances, log)
    manager = PipelineManager(config, utterances, log=log)
    if warps:
        manager.warps = warps

    utts = list(utterances)
    rate_of = {f: m.sample_rate for f, m in manager.audio_metadata.items()}
    rates = sorted(set(rate_of[u.audio_file] for u in utts))
    groups = [[u for u in utts if rate_of[u.audio_file] == rate]
              for rate in rates]
    by_speaker = 'cmvn' in config and config['cmvn']['by_speaker']
    spans = len(groups) > 1 and by_speaker and any(
        len({g for g, gu in enumerate(groups)
             if any(u.speaker == spk for u in gu)}) > 1
        for spk in {u.speaker for u in utts})
    out = {}
    if spans:
        # a speaker has utterances at several sample rates: its CMVN statistics
        # are pooled over all of them (pipeline.py:541-557 accumulates per
        # speaker whatever the rate; test/test_pipeline.py:347-420)
        loaded = [(gu, _load_audios(manager, gu, njobs)) for gu in groups]
        out = _extract_groups_pooled_speakers(manager, loaded, log)
    else:
        for gutts in groups:
            out.update(_extract_group_streamed(
                manager, gutts, log, njobs, gather))
    return FeaturesCollection(
        (u.name, out[u.name]) for u in utts if u.name in out)


def _load_audios(manager, utts, njobs):
    if njobs > 1 and len(utts) > 1:
        with concurrent.futures.ThreadPoolExecutor(njobs) as pool:
            return list(pool.map(manager.get_audio, utts))
    return [manager.get_audio(u) for u in utts]


def _extract_group_streamed(manager, utts, log, njobs, gather=True):
    """Features of utterances that share a sample rate, streamed through the
    fused pipeline in bounded memory (sharded over the ranks of
    torch.distributed when initialised)"""
    config = manager.config
    proc = manager.get_features_processor(utts[0])
    delta = manager.get_delta_processor() if 'delta' in config else None
    cmvn_mode, vad, energy = None, None, None
    if 'cmvn' in config:
        cmvn_mode = 'speaker' if config['cmvn']['by_speaker'] else 'utterance'
        if config['cmvn']['with_vad']:
            vad = manager.get_vad_processor()
            energy = manager.get_energy_processor(utts[0])
    pitch = None
    if 'pitch' in config:
        pitch = (manager.get_pitch_processor(utts[0]),
                 manager.get_pitch_post_processor())
        pitch[1]._validate(2)
    by_speaker = cmvn_mode == 'speaker'
    if by_speaker:    # whole speakers are contiguous (blocks, shards)
        order = np.argsort(np.asarray([u.speaker for u in utts]),
                           kind='stable')
        utts = [utts[i] for i in order]
    items, lengths, int16 = stream.audio_items(utts)
    if vad is not None and not int16:
        # energy keeps the raw scale of float audio (energy.py:158): VAD on
        # non-int16 audio goes through the per-utterance API
        return _extract_group(manager, utts, _load_audios(
            manager, utts, njobs), log)
    has_warp = bool(manager.warps) and proc.name != 'spectrogram'
    warp_of = [manager.get_warp(u) for u in utts] if has_warp else None
    warps = np.asarray(warp_of, np.float32) if has_warp else None
    speakers = [u.speaker for u in utts] if by_speaker else None
    pipe = FusedPipeline(proc, delta=delta, cmvn=cmvn_mode, norm_vars=True,
                         vad=vad, energy=energy, pitch=pitch)
    data, parts = stream.extract_corpus(
        pipe, items, lengths, speakers=speakers, warps=warps, njobs=njobs,
        gather=gather)
    result = {}
    for index, plan, row0, stats, group in parts:
        if stats is not None:
            counts = stats[:, 0, -1]
            if len(counts) and counts.min() < 1.0:
                raise ValueError(
                    'insufficient accumulation of stats for CMVN, '
                    'must be >= 1.0 but is {}'.format(counts.min()))
        for j, i in enumerate(index):
            utt = utts[i]
            a = row0 + int(plan.foffs[j])
            block = data[a:a + int(plan.valid[j])]
            st = None
            if stats is not None:
                st = stats[group[j] if group is not None else j]
            warp = warp_of[i] if warp_of is not None else 1.0
            result[utt.name] = Features._deferred(
                block, _Times(proc, block.shape[0]),
                _Properties(manager, proc, delta, pitch, cmvn_mode, utt,
                            warp, st))
    return result


class _Times:
    """frame timestamps of an utterance, built at first access"""
    __slots__ = ('proc', 'nframes')

    def __init__(self, proc, nframes):
        self.proc, self.nframes = proc, nframes

    def __call__(self):
        return self.proc.times(self.nframes)


class _Properties:
    """properties of an utterance's features with the reference's layout
    (pipeline.py:570-648), built at first access"""
    __slots__ = ('args',)

    def __init__(self, *args):
        self.args = args

    def __call__(self):
        manager, proc, delta, pitch, cmvn_mode, utt, warp, stats = self.args
        props = (proc.get_properties() if proc.name == 'spectrogram'
                 else proc.get_properties(vtln_warp=warp))
        if utt.speaker:
            props['speaker'] = utt.speaker
        props['audio'] = {
            'file': os.path.abspath(utt.audio_file),
            'sample_rate': manager.audio_metadata[utt.audio_file].sample_rate}
        if utt.tstart is not None:
            props['audio']['tstart'] = utt.tstart
            props['audio']['tstop'] = utt.tstop
        props['audio']['duration'] = utt.duration
        carrier = _Carrier(props, proc.ndims)
        if cmvn_mode is not None:
            cmvn = manager.get_cmvn_processor()
            cmvn.add_stats(stats)
            carrier = _Carrier(cmvn.get_properties(carrier), proc.ndims)
        if delta is not None:
            carrier = _Carrier(delta.get_properties(carrier),
                               proc.ndims * (delta.order + 1))
        props = carrier.properties
        if pitch is not None:
            pprops = pitch[1].get_properties(
                _Carrier(pitch[0].get_properties(), 2))
            props.update(
                {k: v for k, v in pprops.items() if k != 'pipeline'})
            for entry in pprops['pipeline']:
                entry['columns'] = [
                    c + carrier.ndims for c in entry['columns']]
                props['pipeline'].append(entry)
        return props


def _check_environment(njobs, log=get_logger('pipeline', 'warning')):
    """Warns when several host jobs meet implicit (OpenMP/BLAS) threading
    (shennong/pipeline.py:299-312); here `njobs` only loads audio files, the
    warning is kept for scripts that rely on it"""
    if njobs == 1:
        return
    nthreads = os.environ.get('OMP_NUM_THREADS')
    if nthreads is None or not nthreads.isdigit() or int(nthreads) != 1:
        log.warning(
            'working on %s threads but implicit parallelism is active, '
            'this may slow down the processing. Set the environment variable '
            'OMP_NUM_THREADS=1 to disable this warning', njobs)


def _warp_setup(configuration, utterances, njobs, log):
    """Shared front end of the warp extractions: (manager, utts, audios)"""
    njobs = get_njobs(njobs, log=log)
    config = _init_config(configuration, log=log)
    manager = PipelineManager(config, utterances, log=log)
    utts = list(utterances)
    if njobs > 1 and len(utts) > 1:
        with concurrent.futures.ThreadPoolExecutor(njobs) as pool:
            audios = list(pool.map(manager.get_audio, utts))
    else:
        audios = [manager.get_audio(u) for u in utts]
    for audio in audios:
        if audio.nchannels != 1:
            raise ValueError(
                'signal must have one dimension, but it has {}'.format(
                    audio.nchannels))
    return manager, utts, audios


def extract_features_warp(configuration, utterances, warp,
                          log=get_logger('pipeline', 'warning'), njobs=1):
    """Features extraction when all features are warped by the same factor

    Mirrors shennong/pipeline.py:669-696 (used by VTLN training): main
    features with ``vtln_warp=warp`` then deltas -- no CMVN, no pitch.
    """
    return extract_features_warp_sweep(
        configuration, utterances, [warp], log=log, njobs=njobs)[float(warp)]


def extract_features_warp_sweep(configuration, utterances, warps,
                                log=get_logger('pipeline', 'warning'),
                                njobs=1):
    """`extract_features_warp` for a whole grid of warps in ONE fused batch

    The VTLN trainer of the reference re-extracts every utterance for every
    warp of its grid (processor/vtln.py:580-622), i.e. calls
    ``extract_features_warp`` ~21 times.  Here the PCM is packed and uploaded
    once and the grid becomes ``len(warps) x len(utterances)`` virtual
    utterances of one batch (same sample ranges, one mel table per warp).

    Returns a dict {warp: FeaturesCollection}.
    """
    warps = [float(w) for w in warps]
    manager, utts, audios = _warp_setup(configuration, utterances, njobs, log)
    if manager.features == 'spectrogram':
        raise ValueError('spectrogram features cannot be warped')
    out = {w: {} for w in warps}
    rates = sorted(set(a.sample_rate for a in audios))
    for rate in rates:
        idx = [i for i, a in enumerate(audios) if a.sample_rate == rate]
        gutts = [utts[i] for i in idx]
        proc = manager.get_features_processor(gutts[0])
        delta = (manager.get_delta_processor()
                 if 'delta' in manager.config else None)
        packed = engine.PackedAudio(
            [audios[i].astype(np.int16).data for i in idx])
        nutt, nwarp = len(idx), len(warps)
        virtual = engine.PackedAudio.from_packed(
            None, np.tile(packed.starts, nwarp),
            np.tile(packed.lengths, nwarp), dev=packed.dev)
        pipe = FusedPipeline(proc, delta=delta)
        dev, offs, _, _ = pipe.run_device(
            virtual, warps=np.repeat(np.asarray(warps, np.float32), nutt))
        data = engine.to_host(dev)
        for k, warp in enumerate(warps):
            for j, utt in enumerate(gutts):
                row = k * nutt + j
                block = data[offs[row]:offs[row + 1]]
                carrier = _Carrier(
                    proc.get_properties(vtln_warp=warp), proc.ndims)
                if delta is not None:
                    carrier = _Carrier(delta.get_properties(carrier),
                                       proc.ndims * (delta.order + 1))
                out[warp][utt.name] = Features(
                    block, proc.times(block.shape[0]), carrier.properties)
    return {w: FeaturesCollection((u.name, out[w][u.name]) for u in utts)
            for w in warps}


def _extract_group(manager, utts, audios, log):
    config = manager.config
    proc = manager.get_features_processor(utts[0])
    for audio in audios:
        if audio.nchannels != 1:
            raise ValueError(
                'signal must have one dimension, but it has {}'.format(
                    audio.nchannels))
    delta = manager.get_delta_processor() if 'delta' in config else None
    cmvn_mode, vad, energy = None, None, None
    if 'cmvn' in config:
        cmvn_mode = 'speaker' if config['cmvn']['by_speaker'] else 'utterance'
        if config['cmvn']['with_vad']:
            vad = manager.get_vad_processor()
            energy = manager.get_energy_processor(utts[0])
    pitch = None
    if 'pitch' in config:
        pitch = (manager.get_pitch_processor(utts[0]),
                 manager.get_pitch_post_processor())
        pitch[1]._validate(2)
    has_warp = bool(manager.warps) and proc.name != 'spectrogram'
    warps = ([manager.get_warp(u) for u in utts] if has_warp else None)
    speakers = [u.speaker for u in utts]

    # energy keeps the raw scale of float audio (energy.py:158): the fused
    # path feeds int16 PCM to every processor, so VAD on non-int16 audio goes
    # through the per-utterance API instead
    raw_int16 = all(a.dtype == np.int16 for a in audios)
    pcms = [a.astype(np.int16).data for a in audios]
    packed = engine.PackedAudio(pcms)

    def run(with_pitch, with_vad):
        pipe = FusedPipeline(
            proc, delta=delta, cmvn=cmvn_mode, norm_vars=True,
            vad=vad if with_vad else None,
            energy=energy if with_vad else None,
            pitch=pitch if with_pitch else None)
        dev, offs, stats, _ = pipe.run_device(
            packed, speakers=speakers, warps=warps)
        return pipe, engine.to_host(dev), offs, (
            engine.to_host(stats) if stats is not None else None)

    weights = None
    if vad is not None and not raw_int16:
        # rare path: per-utterance VAD weights from the host API
        weights = []
        for audio in audios:
            v = vad.process(energy.process(audio))
            weights.append(v.data.reshape(-1).astype(np.float32))
    if weights is not None:
        data, offs, stats, pipe = _run_with_weights(
            proc, delta, cmvn_mode, packed, speakers, warps, weights)
        pitch_blocks = _separate_pitch(pitch, audios) if pitch else None
    else:
        try:
            pipe, data, offs, stats = run(pitch is not None, vad is not None)
            pitch_blocks = None
        except NotImplementedError:
            # features and pitch disagree on the number of frames: paste
            # with the reference's 2-frame tolerance on the host
            pipe, data, offs, stats = run(False, vad is not None)
            pitch_blocks = _separate_pitch(pitch, audios)

    names = getattr(pipe, '_group_names', None)
    return _assemble(manager, proc, delta, pitch, cmvn_mode, utts, audios,
                     warps, data, offs, stats, names, pitch_blocks, log)


def _assemble(manager, proc, delta, pitch, cmvn_mode, utts, audios, warps,
              data, offs, stats, names, pitch_blocks, log):
    """Wraps the rows of a group in Features with the reference's properties
    layout (pipeline.py:570-648)"""
    result = {}
    for i, (utt, audio) in enumerate(zip(utts, audios)):
        block = data[offs[i]:offs[i + 1]]
        warp = warps[i] if warps is not None else 1.0
        props = (proc.get_properties() if proc.name == 'spectrogram'
                 else proc.get_properties(vtln_warp=warp))
        if utt.speaker:
            props['speaker'] = utt.speaker
        props['audio'] = {
            'file': os.path.abspath(utt.audio_file),
            'sample_rate': manager.audio_metadata[utt.audio_file].sample_rate}
        if utt.tstart is not None:
            props['audio']['tstart'] = utt.tstart
            props['audio']['tstop'] = utt.tstop
        props['audio']['duration'] = utt.duration
        carrier = _Carrier(props, proc.ndims)
        if cmvn_mode is not None:
            cmvn = manager.get_cmvn_processor()
            group = names.index(utt.speaker) if names is not None else i
            cmvn.add_stats(stats[group])
            if cmvn.count < 1.0:      # (whatever the utterance's length:
                # the reference raises in cmvn.process, cmvn.py:204-214)
                raise ValueError(
                    'insufficient accumulation of stats for CMVN, '
                    'must be >= 1.0 but is {}'.format(cmvn.count))
            carrier = _Carrier(cmvn.get_properties(carrier), proc.ndims)
        if delta is not None:
            carrier = _Carrier(delta.get_properties(carrier),
                               proc.ndims * (delta.order + 1))
        feat_dim = carrier.ndims
        if pitch is not None and pitch_blocks is None:
            pprops = pitch[1].get_properties(
                _Carrier(pitch[0].get_properties(), 2))
            props = carrier.properties
            props.update(
                {k: v for k, v in pprops.items() if k != 'pipeline'})
            for entry in pprops['pipeline']:
                entry['columns'] = [c + feat_dim for c in entry['columns']]
                props['pipeline'].append(entry)
            feats = Features(block, proc.times(block.shape[0]), props)
        else:
            feats = Features(
                block[:, :feat_dim], proc.times(block.shape[0]),
                carrier.properties)
            if pitch_blocks is not None:
                feats = feats.concatenate(pitch_blocks[i], tolerance=2,
                                          log=log)
        result[utt.name] = feats
    return result


def _extract_groups_pooled_speakers(manager, groups, log):
    """`_extract_group` for several sample-rate groups that share speakers:
    base features and per-utterance statistics of every group first (kept on
    the device: with dither a recomputation would not see the same noise),
    statistics pooled per speaker on the host, then normalisation + deltas
    per group.  Pitch is pasted on the host."""
    torch = engine.require_cuda()
    config = manager.config
    delta = manager.get_delta_processor() if 'delta' in config else None
    with_vad = config['cmvn']['with_vad']
    staged, pooled = [], {}
    for utts, audios in groups:
        for audio in audios:
            if audio.nchannels != 1:
                raise ValueError(
                    'signal must have one dimension, but it has {}'.format(
                        audio.nchannels))
        proc = manager.get_features_processor(utts[0])
        has_warp = bool(manager.warps) and proc.name != 'spectrogram'
        warps = [manager.get_warp(u) for u in utts] if has_warp else None
        pipe = FusedPipeline(proc, delta=delta, cmvn=None)
        plans = pipe._plans()
        packed = engine.PackedAudio([a.astype(np.int16).data for a in audios])
        batch = engine.Batch(plans['feat'], packed, warps)
        layout = engine.RowLayout(batch=batch)
        seed = engine.next_seed() if proc.dither != 0 else 0
        base = engine.compute_features(plans['feat'], batch, seed=seed)
        weights = None
        if with_vad:
            vad = manager.get_vad_processor()
            energy = manager.get_energy_processor(utts[0])
            weights = engine.from_host(np.concatenate([
                vad.process(energy.process(a)).data.reshape(-1)
                for a in audios]).astype(np.float32), np.float32)
        stats = engine.to_host(engine.cmvn_accumulate(base, layout, weights))
        for utt, st in zip(utts, stats):
            pooled[utt.speaker] = pooled.get(utt.speaker, 0) + st
        staged.append((utts, audios, proc, warps, batch, layout, base))
    out = {}
    for utts, audios, proc, warps, batch, layout, base in staged:
        names = sorted({u.speaker for u in utts})
        stats = np.stack([pooled[n] for n in names])
        group = np.array([names.index(u.speaker) for u in utts], np.int32)
        norm = engine.cmvn_norm(engine.from_host(stats, np.float64), True, False)
        data = engine.to_host(engine.deltas(
            base, layout, delta.order if delta is not None else 0,
            delta.window if delta is not None else 1, norm=norm,
            utt_group=torch.from_numpy(group).to('cuda')))
        pitch, pitch_blocks = None, None
        if 'pitch' in config:
            pitch = (manager.get_pitch_processor(utts[0]),
                     manager.get_pitch_post_processor())
            pitch[1]._validate(2)
            pitch_blocks = _separate_pitch(pitch, audios)
        out.update(_assemble(
            manager, proc, delta, pitch, 'speaker', utts, audios, warps, data,
            batch.frame_offsets, stats, names, pitch_blocks, log))
    return out


def _separate_pitch(pitch, audios):
    raw = pitch[0]._process_batch(audios)
    return [pitch[1].process(r) for r in raw]


def _run_with_weights(proc, delta, cmvn_mode, packed, speakers, warps,
                      weights):
    """CMVN weighted by host-provided VAD decisions"""
    torch = engine.require_cuda()
    pipe = FusedPipeline(proc, delta=delta, cmvn=None)
    plans = pipe._plans()
    batch = engine.Batch(plans['feat'], packed, warps)
    layout = engine.RowLayout(batch=batch)
    seed = engine.next_seed() if proc.dither != 0 else 0
    base = engine.compute_features(plans['feat'], batch, seed=seed)
    w = engine.from_host(np.concatenate(weights), np.float32)
    stats = engine.cmvn_accumulate(base, layout, w)
    utt_group = None
    if cmvn_mode == 'speaker':
        names = sorted(set(speakers))
        group = np.array([names.index(s) for s in speakers], dtype=np.int64)
        ptr = np.concatenate(
            ([0], np.cumsum(np.bincount(group, minlength=len(names)))))
        stats = engine.cmvn_reduce_groups(
            stats, ptr, np.argsort(group, kind='stable'), len(names))
        utt_group = torch.from_numpy(group.astype(np.int32)).to('cuda')
        pipe._group_names = names
    norm = engine.cmvn_norm(stats, True, False)
    order = delta.order if delta is not None else 0
    window = delta.window if delta is not None else 1
    out = engine.deltas(base, layout, order, window, norm=norm,
                        utt_group=utt_group)
    return (engine.to_host(out), batch.frame_offsets, engine.to_host(stats),
            pipe)
