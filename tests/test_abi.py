"""The C-ABI library loads and exports every symbol include/nsw.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'nsw.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nsw_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ('nsw_iaf_create', 'nsw_iaf_forward_host', 'nsw_iaf_forward_device',
              'nsw_fastgen_create', 'nsw_fastgen_run_host', 'nsw_last_error'):
        assert s in syms


def test_library_builds_loads_and_exports_all_symbols():
    from nsynth_wavenet_b200 import _lib
    lib = _lib.load()
    assert os.path.exists(_lib.lib_path())
    raw = ctypes.CDLL(_lib.lib_path())
    for s in declared_symbols():
        assert hasattr(raw, s), 'libnsw_b200.so does not export ' + s
    assert set(declared_symbols()) == set(_lib.PROTOTYPES), 'ctypes prototypes out of sync'
    assert lib.nsw_version() >= 100


def test_struct_sizes_match_header():
    from nsynth_wavenet_b200 import _lib
    assert ctypes.sizeof(_lib.nsw_iaf_config) == 4 * (1 + 8 + 6 + 4 + 4 + 5)
    assert ctypes.sizeof(_lib.nsw_wavenet_config) == 4 * (10 + 4 + 4 + 4)
    assert ctypes.sizeof(_lib.nsw_tensor) == 8 + 8 + 8 + 32


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'nsynth_wavenet_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'oracle' not in src.replace('oracle/', '').lower() or f == 'none', (dp, f)
