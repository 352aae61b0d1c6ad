"""dsp_final_b200 -- B200 (sm_100a) implementation of the src/dsp -> MFCC-retrieval
hot path of Audiofool934/dsp-final, behind the reference's own Python signatures.

    dsp_final_b200.dsp        drop-in for the reference's src/dsp (fft, stft, mfcc modules)
    dsp_final_b200.retrieval  drop-in for src/retrieval/retrieval.py's scoring functions
    dsp_final_b200.batch      batched device / host entry points (the throughput API)
    dsp_final_b200.cache      batched writer of the reference's .npy feature-cache format
    dsp_final_b200.dist       clip sharding + the one all-gather of database embeddings

All arithmetic runs in hand-written CUDA kernels inside _native/libdspx.so
(C ABI: include/dspx.h).  Importing this package does not initialise CUDA, so it
is safe before a fork (DataLoader workers); the first compute call does.
"""
from .dsp.mfcc import MfccConfig  # noqa: F401

__all__ = ["MfccConfig"]
__version__ = "0.1.0"
