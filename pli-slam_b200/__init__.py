"""plf-b200: B200-native stereo point-line frontend (drop-in for PLI-SLAM's ORB / LSD+LBD / stereo-matching path).

The product is `libplf_b200.so` (hand-written sm_100a CUDA behind the C ABI of include/plf_b200.h).  This package is
the thin Python host mirror used by tests and bench.py; the C++ host mirror of the reference classes lives in
`pli-slam_b200/host/`.  There is no CPU fallback: `load_product()` raises when the CUDA library is absent.
"""
from .binding import (Frontend, Library, Params, PlfError, BatchResult, load_product, load_oracle, KEYPOINT_DT,
                      KEYLINE_DT, PROJ_QUERY_DT, FRAME_QUERY_DT, TRACK_LINE_DT, ABI_SYMBOLS, PRODUCT_LIB, ORACLE_LIB)
from .synth import synth_pair, synth_batch, synth_curvy, synth_vocabulary, rectify_maps, EUROC_CALIB
from .pool import shard_streams, stream_owner, stream_seed, DevicePool
