from nsynth_wavenet_b200.auxilaries.mel_extractor import *  # noqa: F401,F403
from nsynth_wavenet_b200.auxilaries.mel_extractor import (  # noqa: F401
    batch_melspectrogram, melspectrogram, mel_params)
