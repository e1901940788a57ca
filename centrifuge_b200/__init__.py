"""centrifuge_b200 -- B200-native (sm_100a) Centrifuge classification hot path.

The product is the native library (csrc/, include/cfb200.h); this package only holds the build
recipe and a ctypes binding used by tests and bench.py.
"""
__version__ = "0.1"
