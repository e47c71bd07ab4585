"""emri_frequencydomainwaveforms_b200 -- B200-native FD EMRI mode-sum + likelihood hot path.

Drop-in for the path the reference scripts drive through FastEMRIWaveforms:
``GenerateEMRIWaveform(..., sum_kwargs={'output_type': 'fd'})`` / ``FDInterpolatedModeSum`` and the
lisatools ``inner_product`` / ``Likelihood`` interface.  All arithmetic on the path runs in
hand-written sm_100a CUDA kernels behind the C-ABI in ``include/emrifd.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"
