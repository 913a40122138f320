"""Import shim for `from auxilaries import utils, mel_extractor` (eval_*.py:7)."""
