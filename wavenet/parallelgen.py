from nsynth_wavenet_b200.wavenet.parallelgen import *  # noqa: F401,F403
from nsynth_wavenet_b200.wavenet.parallelgen import (  # noqa: F401
    get_default_shadow_dict, load_parallelgen, synthesis)
