from nsynth_wavenet_b200.wavenet.fastgen import *  # noqa: F401,F403
from nsynth_wavenet_b200.wavenet.fastgen import (  # noqa: F401
    calculate_cond_vars, encode, get_ema_shadow_dict, load_batch, load_cond_layers, load_deconv_stack,
    load_fastgen, save_batch, synthesis)
