"""Small host utilities (counterpart of shennong/utils.py)"""

import multiprocessing

import numpy as np

from shennong_b200.logger import null_logger


def get_njobs(njobs=None, log=null_logger()):
    """Number of parallel jobs, clipped to the number of CPU cores

    On the GPU engine `njobs` only sizes the host-side audio loading pool:
    the extraction itself is one batched launch.  Raises ValueError when
    `njobs` is not strictly positive (shennong/utils.py:16-47).
    """
    ncores = multiprocessing.cpu_count()
    if njobs is None:
        return ncores
    if njobs <= 0:
        raise ValueError(
            'njobs must be strictly positive, it is {}'.format(njobs))
    if njobs > ncores:
        log.warning(
            'asking %d CPU cores but reducing to %d (max available)',
            njobs, ncores)
        return ncores
    return njobs


def array2list(seq):
    """`seq` with every numpy array (also inside nested dicts) turned into a list"""
    if isinstance(seq, np.ndarray):
        return seq.tolist()
    if isinstance(seq, dict):
        return {key: array2list(value) for key, value in seq.items()}
    return seq


def list2array(seq):
    """Inverse of :func:`array2list`: lists (also inside nested dicts) become arrays"""
    if isinstance(seq, dict):
        return {key: list2array(value) for key, value in seq.items()}
    return np.asarray(seq) if isinstance(seq, list) else seq


def dict_equal(dict1, dict2):
    """Equality of dicts that may hold numpy arrays (shennong/utils.py:78-96)"""
    return array2list(dict1) == array2list(dict2)


def list_files_with_extension(directory, extension, abspath=False,
                              realpath=True, recursive=True):
    """Sorted list of the files of `directory` whose name ends with
    `extension` (shennong/utils.py:99-147): paths relative to `directory`
    unless `abspath` or `realpath`; `recursive` walks the subdirectories"""
    import os
    found = []
    for root, _, names in os.walk(directory):
        for name in names:
            if name.endswith(extension):
                found.append(os.path.join(root, name))
        if not recursive:
            break
    if abspath:
        found = [os.path.abspath(f) for f in found]
    if realpath:
        found = [os.path.realpath(f) for f in found]
    if not (abspath or realpath):
        found = [os.path.relpath(f, directory) for f in found]
    return sorted(found)


class CatchExceptions:
    """Decorator for command-line entry points: runs the function and turns
    the usual failures into a one-line message on stderr and exit code 1
    (same surface as shennong/utils.py:147-190)"""
    def __init__(self, function):
        self.function = function

    def __call__(self):
        try:
            self.function()
        except KeyboardInterrupt:
            self.exit('keyboard interruption, exiting')
        except (OSError, ValueError, RuntimeError, AssertionError) as err:
            self.exit('fatal error: {}'.format(err))

    @staticmethod
    def exit(message):
        import sys
        sys.stderr.write(message.strip() + '\n')
        sys.exit(1)
