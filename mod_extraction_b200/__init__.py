"""B200-native renderer for the effect hot path of christhetree/mod_extraction.

Public surface (reference module -> drop-in here):
    mod_extraction.fx           -> mod_extraction_b200.fx           MonoFlangerChorusModule, apply_tremolo
    mod_extraction.modulations  -> mod_extraction_b200.modulations  make_mod_signal, make_rand_mod_signal,
                                                                    make_quasi_periodic, make_combined_mod_sig, find_corners
    mod_extraction.util         -> mod_extraction_b200.util         linear_interpolate_last_dim + host RNG helpers
    Spectral2DCNN.spectrogram   -> mod_extraction_b200.models       MelSpectrogram / LogMelSpectrogram
    pedalboard phaser call site -> mod_extraction_b200.phaser       Phaser, apply_pedalboard_phaser (parity unpinned)
    interwoven data path        -> mod_extraction_b200.render       InterwovenRenderer
All arithmetic runs in libmodfx.so (hand-written sm_100a kernels behind the C ABI of include/modfx.h).
"""
__version__ = "0.1.0"
