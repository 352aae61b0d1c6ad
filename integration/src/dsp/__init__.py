# Empty, like the reference's src/dsp/__init__.py: callers import from the submodules
# (src.dsp.fft, src.dsp.stft, src.dsp.mfcc), which are the shims next to this file.
