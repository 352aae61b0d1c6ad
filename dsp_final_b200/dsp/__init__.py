"""Drop-in for the reference package src/dsp (same module and function names)."""
from .fft import fft, ifft, rfft  # noqa: F401
from .mfcc import MfccConfig, dct_type_2, log_mel_spectrogram, mel_filterbank, mfcc  # noqa: F401
from .stft import WindowType, frame_signal, stft  # noqa: F401
