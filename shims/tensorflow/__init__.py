"""Minimal stand-in for the handful of TensorFlow 1.x symbols the reference's eval CLIs touch OUTSIDE the model
(eval_wavenet.py:4,17-23,33-34,59 and eval_parallel_wavenet.py:4,17-23,33-34,59): tf.gfile, tf.logging and
tf.train.latest_checkpoint / checkpoint_exists.  The model itself never imports TensorFlow here -- it runs in
libnsw_b200.so behind wavenet.fastgen / wavenet.parallelgen.

Put this directory LAST on PYTHONPATH so that a real TensorFlow, when installed, wins:

    PYTHONPATH=/path/to/nsynth_wavenet_b200_repo:/path/to/nsynth_wavenet_b200_repo/shims \\
        python eval_parallel_wavenet.py --ckpt_dir ... --source_path ... --save_path ...
"""
import logging as _pylog
import os as _os
from types import SimpleNamespace as _NS

__version__ = '1.x-shim (nsynth_wavenet_b200)'

_log = _pylog.getLogger('tensorflow')
if not _log.handlers:
    _h = _pylog.StreamHandler()
    _h.setFormatter(_pylog.Formatter('%(levelname)s:tensorflow:%(message)s'))
    _log.addHandler(_h)
    _log.setLevel(_pylog.INFO)
    _log.propagate = False

_LEVELS = {'DEBUG': _pylog.DEBUG, 'INFO': _pylog.INFO, 'WARN': _pylog.WARNING, 'WARNING': _pylog.WARNING,
           'ERROR': _pylog.ERROR, 'FATAL': _pylog.CRITICAL}


def _set_verbosity(level):
    """tf.logging.set_verbosity accepts the integer constants; the CLIs pass the --log string (eval_wavenet.py:19)."""
    if isinstance(level, str):
        level = _LEVELS.get(level.upper(), _pylog.INFO)
    _log.setLevel(level)


logging = _NS(DEBUG=_pylog.DEBUG, INFO=_pylog.INFO, WARN=_pylog.WARNING, ERROR=_pylog.ERROR, FATAL=_pylog.CRITICAL,
              debug=_log.debug, info=_log.info, warn=_log.warning, warning=_log.warning, error=_log.error,
              fatal=_log.critical, set_verbosity=_set_verbosity, get_verbosity=lambda: _log.level)

gfile = _NS(IsDirectory=_os.path.isdir, ListDirectory=lambda d: sorted(_os.listdir(d)), Exists=_os.path.exists,
            MakeDirs=lambda d: _os.makedirs(d, exist_ok=True), Glob=lambda p: sorted(__import__('glob').glob(p)))


def _latest_checkpoint(checkpoint_dir, latest_filename=None):
    """tf.train.latest_checkpoint: the prefix named by the directory's `checkpoint` state file (None if absent)."""
    from nsynth_wavenet_b200 import tf_bundle
    return tf_bundle.latest_checkpoint(checkpoint_dir)


def _checkpoint_exists(checkpoint_prefix):
    """tf.train.checkpoint_exists: a V2 bundle (<prefix>.index) or a V1 file with that exact name."""
    if not checkpoint_prefix:
        return False
    from nsynth_wavenet_b200 import tf_bundle
    return tf_bundle.is_bundle_prefix(checkpoint_prefix) or _os.path.isfile(checkpoint_prefix)


train = _NS(latest_checkpoint=_latest_checkpoint, checkpoint_exists=_checkpoint_exists)
