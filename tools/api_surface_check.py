#!/usr/bin/env python
"""Compares the public API surface of the on-path classes of a shennong source
tree (parsed with `ast`, never imported) with shennong_b200: public methods /
attributes of every class, and argument names + defaults of every public
method and constructor.

    python tools/api_surface_check.py /root/reference > profiles/rNN_api_surface_check.txt
"""
import ast
import importlib
import inspect
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CLASSES = {
    'shennong/audio.py': ['Audio'],
    'shennong/features.py': ['Features'],
    'shennong/features_collection.py': ['FeaturesCollection'],
    'shennong/utterances.py': ['Utterance', 'Utterances'],
    'shennong/frames.py': ['Frames'],
    'shennong/base.py': ['BaseProcessor'],
    'shennong/processor/base.py': ['FeaturesProcessor', 'FramesProcessor', 'MelFeaturesProcessor'],
    'shennong/processor/mfcc.py': ['MfccProcessor'],
    'shennong/processor/filterbank.py': ['FilterbankProcessor'],
    'shennong/processor/plp.py': ['PlpProcessor', 'RastaFilter'],
    'shennong/processor/spectrogram.py': ['SpectrogramProcessor'],
    'shennong/processor/energy.py': ['EnergyProcessor'],
    'shennong/processor/pitch_kaldi.py': ['KaldiPitchProcessor', 'KaldiPitchPostProcessor'],
    'shennong/postprocessor/base.py': ['FeaturesPostProcessor'],
    'shennong/postprocessor/cmvn.py': ['CmvnPostProcessor', 'SlidingWindowCmvnPostProcessor'],
    'shennong/postprocessor/delta.py': ['DeltaPostProcessor'],
    'shennong/postprocessor/vad.py': ['VadPostProcessor'],
    'shennong/pipeline_manager.py': ['PipelineManager'],
    'shennong/serializers.py': ['FeaturesSerializer', 'NumpySerializer', 'MatlabSerializer',
                                'PickleSerializer', 'KaldiSerializer', 'CsvSerializer'],
}


FUNCTION_MODULES = ['shennong/pipeline.py', 'shennong/utils.py', 'shennong/serializers.py',
                    'shennong/window.py', 'shennong/logger.py', 'shennong/__init__.py']


def ref_signature(fn):
    names = [a.arg for a in fn.args.args][1:]
    defaults = []
    for d in fn.args.defaults:
        try:
            defaults.append(ast.literal_eval(d))
        except ValueError:
            defaults.append(ast.unparse(d))
    first = len(names) - len(defaults)
    return [(n, defaults[i - first] if i >= first else '<required>') for i, n in enumerate(names)]


def same_default(ref, ours):
    if ref == ours:
        return True
    if isinstance(ref, str) and ('logger' in ref):
        return True                      # a logger object
    try:
        if isinstance(ref, str):
            ref = eval(ref, {'np': __import__('numpy')})
        return abs(float(ref) - float(ours)) < 1e-6
    except Exception:
        return False


def main():
    ref_root = sys.argv[1]
    import shennong_b200.compat as compat
    compat.install(force=True)
    nclasses = nmethods = 0
    problems = []
    for path, classes in CLASSES.items():
        tree = ast.parse(open(os.path.join(ref_root, path)).read())
        mod = importlib.import_module(path[:-3].replace('/', '.'))
        for node in tree.body:
            if not (isinstance(node, ast.ClassDef) and node.name in classes):
                continue
            nclasses += 1
            cls = getattr(mod, node.name, None)
            if cls is None:
                problems.append(f'{node.name}: class missing')
                continue
            for item in node.body:
                if isinstance(item, ast.Assign):
                    for t in item.targets:
                        if isinstance(t, ast.Name) and not t.id.startswith('_') and not hasattr(cls, t.id):
                            problems.append(f'{node.name}.{t.id}: attribute missing')
                if not isinstance(item, ast.FunctionDef):
                    continue
                if item.name.startswith('_') and item.name != '__init__':
                    continue
                nmethods += 1
                if not hasattr(cls, item.name):
                    problems.append(f'{node.name}.{item.name}: missing')
                    continue
                if isinstance(inspect.getattr_static(cls, item.name), property):
                    continue
                try:
                    sig = inspect.signature(getattr(cls, item.name))
                except (TypeError, ValueError):
                    continue
                ours = [(n, p.default if p.default is not inspect.Parameter.empty else '<required>')
                        for n, p in sig.parameters.items()
                        if n != 'self' and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
                ref = ref_signature(item)
                if [n for n, _ in ref] != [n for n, _ in ours]:
                    problems.append(f'{node.name}.{item.name}: arguments {[n for n, _ in ref]} '
                                    f'vs {[n for n, _ in ours]}')
                    continue
                for (n, rd), (_, od) in zip(ref, ours):
                    if rd == '<required>' and od is None:
                        continue         # an argument the engine does not need
                    if not same_default(rd, od):
                        problems.append(f'{node.name}.{item.name}: default of {n}: {rd!r} vs {od!r}')
    nfuncs = 0
    for path in FUNCTION_MODULES:
        tree = ast.parse(open(os.path.join(ref_root, path)).read())
        name = path[:-3].replace('/', '.').replace('.__init__', '')
        mod = importlib.import_module(name)
        for item in tree.body:
            if not isinstance(item, ast.FunctionDef) or item.name.startswith('_'):
                continue
            nfuncs += 1
            fn = getattr(mod, item.name, None)
            if fn is None:
                problems.append(f'{name}.{item.name}: function missing')
                continue
            ours = [(n, p.default if p.default is not inspect.Parameter.empty else '<required>')
                    for n, p in inspect.signature(fn).parameters.items()
                    if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
            names = [a.arg for a in item.args.args]
            defaults = []
            for d in item.args.defaults:
                try:
                    defaults.append(ast.literal_eval(d))
                except ValueError:
                    defaults.append(ast.unparse(d))
            first = len(names) - len(defaults)
            ref = [(n, defaults[i - first] if i >= first else '<required>') for i, n in enumerate(names)]
            if [n for n, _ in ref] != [n for n, _ in ours]:
                problems.append(f'{name}.{item.name}: arguments {[n for n, _ in ref]} vs {[n for n, _ in ours]}')
                continue
            for (n, rd), (_, od) in zip(ref, ours):
                if not same_default(rd, od) and not (isinstance(rd, str) and rd in ('int', 'str', 'tuple')):
                    problems.append(f'{name}.{item.name}: default of {n}: {rd!r} vs {od!r}')
    print(f'# {nfuncs} public module-level functions compared')
    print(f'# {nclasses} classes, {nmethods} public methods / constructors compared '
          f'({ref_root} parsed with ast, shennong_b200 introspected)')
    for p in problems:
        print(p)
    print(f'# {len(problems)} difference(s)')


if __name__ == '__main__':
    main()
