"""B200-native generation stack for the WaveNet / Parallel-WaveNet (IAF) vocoders of
bfs18/nsynth_wavenet: hand-written sm_100a CUDA behind a C ABI (include/nsw.h) and
the reference's own Python entry points (wavenet.parallelgen / wavenet.fastgen)."""
from . import _lib  # noqa: F401
from .engine import IAFEngine, FastgenEngine, TeacherEngine, iaf_config, wavenet_config  # noqa: F401

__all__ = ['IAFEngine', 'FastgenEngine', 'TeacherEngine', 'iaf_config', 'wavenet_config']
