from dsp_final_b200.dsp.fft import _next_pow_two, fft, ifft, rfft  # noqa: F401
