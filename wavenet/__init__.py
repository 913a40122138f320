"""Import shim so that the reference's own `from wavenet import fastgen, parallelgen`
(eval_wavenet.py:6, eval_parallel_wavenet.py:6) resolves to the B200 implementation."""
