from dsp_final_b200.dsp.stft import WindowType, _get_window, frame_signal, stft  # noqa: F401
