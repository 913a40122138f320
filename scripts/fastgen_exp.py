"""Fastgen timing experiment: run the persistent kernel for --steps samples under the
NSW_FASTGEN_FLAGS / NSW_FASTGEN_DEBUG switches given in the environment and print us/step."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=16000)
    ap.add_argument('--flags', type=str, default='0')
    ap.add_argument('--debug', action='store_true')
    a = ap.parse_args()
    import torch
    from argparse import Namespace
    from nsynth_wavenet_b200 import FastgenEngine
    from oracle import wavenet_oracle as O
    with open(os.path.join(ROOT, 'nsynth_wavenet_b200', 'config_jsons', 'wavenet_mol.json')) as f:
        hp = Namespace(**json.load(f))
    w = O.init_teacher_weights(hp, seed=12345)
    eng = FastgenEngine(hp, w, device=0)
    g = torch.Generator(device='cpu').manual_seed(1)
    enc = (torch.rand((1, a.steps, 256), generator=g) * 2 - 1).to('cuda:0')
    ref = None
    for spec in a.flags.split(','):
        # spec = flags[:l2last[:polldelay[:critdelay]]]
        parts = spec.split(':')
        fl = parts[0]
        os.environ.pop('NSW_FASTGEN_GENERIC', None)
        if fl in ('default', 'generic'):   # the library's own defaults (lean compile-time build / run-time-switch build)
            if fl == 'generic':
                os.environ['NSW_FASTGEN_GENERIC'] = '1'
            for k in ('NSW_FASTGEN_FLAGS', 'NSW_FASTGEN_L2LAST', 'NSW_FASTGEN_POLLDELAY', 'NSW_FASTGEN_CRITDELAY'):
                os.environ.pop(k, None)
        else:
            os.environ['NSW_FASTGEN_FLAGS'] = fl
            os.environ['NSW_FASTGEN_L2LAST'] = parts[1] if len(parts) > 1 else '0'
            os.environ['NSW_FASTGEN_POLLDELAY'] = parts[2] if len(parts) > 2 else '0'
            os.environ['NSW_FASTGEN_CRITDELAY'] = parts[3] if len(parts) > 3 else '0'
        fl = spec
        os.environ.pop('NSW_FASTGEN_DEBUG', None)
        eng.run_device(enc[:, :2048], seed=1)
        torch.cuda.synchronize()
        audio = eng.run_device(enc, seed=2)
        torch.cuda.synchronize()
        ms = eng.last_timing()
        au = audio[0] if isinstance(audio, (tuple, list)) else audio
        au = au.float().cpu()
        if ref is None:
            ref = au
        same = bool(torch.equal(ref, au))
        print('flags %14s: %.2f us/step  rtf %.3f  same_audio_as_first=%s' % (
            fl, 1e3 * ms / a.steps, a.steps / (ms * 1e-3) / 16000.0, same), flush=True)
        if a.debug:
            os.environ['NSW_FASTGEN_DEBUG'] = '1'
            eng.run_device(enc[:, :4096], seed=2)
            torch.cuda.synchronize()


if __name__ == '__main__':
    main()
