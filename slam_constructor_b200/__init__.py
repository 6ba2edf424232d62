"""slam-constructor B200 engine: scan scoring, grid update and max-pyramid on sm_100a.

The product is lib/libslamgpu.so (hand-written CUDA behind the C ABI of include/slamgpu.h).
This package is the thin Python host layer over that ABI (ctypes) used by the tests and
bench.py; the C++ plugin adapters that mirror the reference's interfaces live in host/.
There is no CPU fallback: without the library or without a B200 every compute call raises.
"""
from .capi import (  # noqa: F401
    CELL_AFFINE, CELL_CREDIBILIST, CELL_GMAPPING, CELL_LWW, CELL_MEAN, CELL_TBM_CONSISTENT, CELL_TBM_UNKNOWN_EVEN, EST_AREA, EST_CONST,
    GROW_NONE, GROW_PLAIN, GROW_TILED, OIE_DISCREPANCY, OIE_OCCUPANCY, OOPE_GMAPPING, OOPE_MAX, OOPE_MEAN, OOPE_OBSTACLE,
    OOPE_OVERLAP, TRIG_DEVICE, TRIG_HOST, Context, Estimator, GridMap, Particles, Pyramid, Scan, SlamGpuError, SpeParams, GmCache, estimator, lib,
    library_path, spe_params, point_weights, mapping_quality, SPW_EVEN, SPW_VINY, SPW_AHR, OMQE_IDLE, OMQE_AHR)
