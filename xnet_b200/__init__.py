"""xnet_b200: B200-native batched per-zone implicit step for XNet (hand-written sm_100a CUDA behind
a C ABI, include/xnet_b200.h).  This package is the thin host-side mirror used by tests and
bench.py; the product is libxnet_b200.so.  There is no CPU fallback: without the built library or
a CUDA device every compute call raises."""
from .api import (Network, Context, Controls, Fixture, XnetB200Error, lib, lib_path, read_controls,
                  build_zone_set, load_fixture)
from .ts_io import write_ts, write_ev, read_ts, compare_ts, name_ordered

__all__ = ['Network', 'Context', 'Controls', 'Fixture', 'XnetB200Error', 'lib', 'lib_path',
           'read_controls', 'build_zone_set', 'load_fixture', 'write_ts', 'write_ev', 'read_ts', 'compare_ts', 'name_ordered']
