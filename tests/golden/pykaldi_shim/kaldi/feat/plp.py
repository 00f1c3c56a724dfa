from .mel import MelBanksOptions
from .window import FrameExtractionOptions


class PlpOptions:
    def __init__(self):
        self.frame_opts = FrameExtractionOptions()
        self.mel_opts = MelBanksOptions()
        self.lpc_order = 12
        self.num_ceps = 13
        self.use_energy = True
        self.energy_floor = 0.0
        self.raw_energy = True
        self.compress_factor = 0.33333
        self.cepstral_lifter = 22
        self.cepstral_scale = 1.0
        self.htk_compat = False
