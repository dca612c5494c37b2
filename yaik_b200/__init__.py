"""yaik_b200 — B200-native encoder-analysis stage of the YAIK image codec (alpha-zero tile rejection, the 7-pass
gradient tile cascade, 8x8 range compression) behind a C ABI (include/yaik_b200.h).

The compute path is CUDA only (yaik_b200/csrc, sm_100a).  Importing the package does not load the library;
`yaik_b200.capi.load_library()` does and raises if it has not been built."""
__all__ = ["capi", "synth"]
