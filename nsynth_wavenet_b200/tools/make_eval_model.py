"""tools/make_eval_model.py of the reference (tools/make_eval_model.py:8-34) without TensorFlow: strip a training
checkpoint directory to the variables generation needs -- the ExponentialMovingAverage shadows, under their shadow
names -- and write them as a fresh TF-V2 bundle plus the `checkpoint` state file and the config json.

    python -m nsynth_wavenet_b200.tools.make_eval_model --ckpt_dir logs/ns_pwn --save_dir ns_pwn-eval
"""
from __future__ import annotations

import argparse
import glob
import os
import shutil

from .. import tf_bundle


def save_eval_model(ckpt_dir, save_dir):
    if os.path.exists(save_dir):
        shutil.rmtree(save_dir)                                        # make_eval_model.py:9-11
    os.mkdir(save_dir)
    prefix = tf_bundle.latest_checkpoint(ckpt_dir)                      # tf.train.get_checkpoint_state (:13)
    if prefix is None:
        raise FileNotFoundError('no TF-V2 checkpoint bundle in {!r}'.format(ckpt_dir))
    ema = tf_bundle.read_bundle(prefix, names=lambda n: 'ExponentialMovingAverage' in n)   # :17-19
    if not ema:
        raise ValueError('{} holds no ExponentialMovingAverage variables'.format(prefix))
    base = os.path.basename(prefix)
    tf_bundle.write_bundle(os.path.join(save_dir, base), ema)           # Saver(var_list=eval_vars).save (:25-27)
    with open(os.path.join(save_dir, 'checkpoint'), 'wt', encoding='utf-8') as f:
        f.write('model_checkpoint_path: "{}"'.format(base))             # :29-31
    for json_file in glob.glob(os.path.join(ckpt_dir, '*.json'))[:1]:
        shutil.copy(json_file, save_dir)                                # :33-34
    return os.path.join(save_dir, base)


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--ckpt_dir', required=True)
    parser.add_argument('--save_dir', required=True)
    args = parser.parse_args()
    save_eval_model(args.ckpt_dir, args.save_dir)
