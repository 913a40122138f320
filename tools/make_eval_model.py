"""Same command line as the reference's tools/make_eval_model.py."""
from nsynth_wavenet_b200.tools.make_eval_model import save_eval_model  # noqa: F401

if __name__ == '__main__':
    import runpy
    runpy.run_module('nsynth_wavenet_b200.tools.make_eval_model', run_name='__main__')
