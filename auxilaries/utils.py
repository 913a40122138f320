from nsynth_wavenet_b200.auxilaries.utils import *  # noqa: F401,F403
