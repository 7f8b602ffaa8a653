"""CPU: the C-ABI library loads and exports every symbol include/sis_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from synthesis_in_style_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'sis_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sis_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported():
    names = declared_symbols()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/sis_b200.h but not exported by libsis_b200.so'


def test_python_binding_covers_header():
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()
    lib = _lib.load()
    assert lib.sis_version() >= 100
    assert lib.sis_upfirdn2d_out_size(17, 1, 1, 1, 1, 4) == 16      # Blur after the up-conv: (2H+1) -> 2H
    assert lib.sis_upfirdn2d_out_size(8, 2, 1, 2, 1, 4) == 16       # ToRGB skip upsample
    assert lib.sis_upfirdn2d_out_size(16, 1, 2, 1, 1, 4) == 8


def test_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    # argument validation happens before any CUDA call
    st = lib.sis_fused_bias_act(None, None, None, None, 0, 16, 0, 0, 3, 0, 0.2, 1.4, None)
    assert st != 0 and b'fused_bias_act' in lib.sis_last_error()
    st = lib.sis_upfirdn2d(None, None, None, 0, 1, 4, 4, 1, 40, 40, 1, 1, 1, 1, 0, 0, 0, 0, None)
    assert st != 0 and b'kernel must be between' in lib.sis_last_error()
    h = ctypes.c_void_p()
    assert lib.sis_generator_create(100, 512, 8, 2, ctypes.byref(h)) != 0
    assert lib.sis_generator_create(32, 64, 2, 2, ctypes.byref(h)) == 0
    assert lib.sis_generator_n_latent(h) == 8 and lib.sis_generator_num_layers(h) == 7
    c, r = ctypes.c_int(), ctypes.c_int()
    assert lib.sis_generator_activation_shape(h, 7, ctypes.byref(c), ctypes.byref(r)) == 0
    assert (c.value, r.value) == (512, 32)
    args = _lib.ForwardArgs()
    args.batch = 1
    assert lib.sis_generator_forward(h, ctypes.byref(args), None) != 0     # not prepared
    assert b'prepare' in lib.sis_last_error()
    assert lib.sis_generator_destroy(h) == 0
