"""Drop-in for the reference package src/dsp: submodules fft, stft, mfcc with the same names.

Like the reference's (empty) src/dsp/__init__.py this does not re-export the functions: `fft`,
`stft` and `mfcc` are both submodule and function names, and re-exporting them would shadow the
submodules for `import dsp_final_b200.dsp.mfcc as m`.
"""
