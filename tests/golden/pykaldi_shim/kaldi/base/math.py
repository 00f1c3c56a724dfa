import math as _math


def log(x):
    """kaldi::Log(double)"""
    return _math.log(float(x))
