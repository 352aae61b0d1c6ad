from dsp_final_b200.dsp.mfcc import (  # noqa: F401
    MfccConfig, _dct_basis, _hz_to_mel, _mel_filterbank_cached, _mel_to_hz, dct_type_2, log_mel_spectrogram,
    mel_filterbank, mfcc,
)
