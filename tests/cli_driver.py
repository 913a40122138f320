"""TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so the reference's eval CLIs cannot be executed
there.  This is the call sequence of eval_wavenet.py:10-69 / eval_parallel_wavenet.py:10-67 written out again
(same tf.* calls through the shim, same module functions, same argument flow) so that the drop-in modules can be driven
end to end with real engines on the GPU; tests/test_cli_dropin.py runs the UNMODIFIED reference scripts wherever
/root/reference is present."""
import glob
import json
import os
from argparse import Namespace

import tensorflow as tf  # the shim (shims/tensorflow) unless a real TensorFlow is installed

from auxilaries import mel_extractor, utils
from wavenet import fastgen, parallelgen


def generate(kind, ckpt_dir, source_path, save_path, sample_length=-1, batch_size=1, log='INFO'):
    source_path = utils.shell_path(source_path)
    ckpt_dir = utils.shell_path(ckpt_dir)
    save_path = utils.shell_path(save_path)
    if not os.path.exists(save_path):
        os.mkdir(save_path)
    tf.logging.set_verbosity(log)
    assert tf.gfile.IsDirectory(ckpt_dir)
    checkpoint_path = tf.train.latest_checkpoint(ckpt_dir)
    assert tf.train.checkpoint_exists(checkpoint_path)
    json_in_dir = glob.glob(os.path.join(ckpt_dir, '*.json'))
    assert len(json_in_dir) == 1
    with open(json_in_dir[0], 'rt') as f:
        hparams = Namespace(**json.load(f))
    assert tf.gfile.IsDirectory(source_path)
    files = sorted(os.path.join(source_path, f) for f in tf.gfile.ListDirectory(source_path)
                   if f.lower().endswith('.wav'))
    for start in range(0, len(files), batch_size):
        tf.logging.info('generating batch {:d}'.format(start // batch_size))
        batch_files = files[start:start + batch_size]
        save_names = [os.path.join(save_path, 'gen_' + os.path.splitext(os.path.basename(f))[0] + '.wav')
                      for f in batch_files]
        batch_data = fastgen.load_batch(batch_files, sample_length=sample_length)
        if kind == 'wavenet':
            encoding = fastgen.encode(hparams, batch_data, checkpoint_path)
            fastgen.synthesis(hparams, encoding, save_names, checkpoint_path)
        else:
            mel_data = mel_extractor.batch_melspectrogram(batch_data)
            parallelgen.synthesis(hparams, mel_data, save_names, checkpoint_path)
    return files
