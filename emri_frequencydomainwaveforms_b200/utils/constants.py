"""Physical constants used on the FD hot path.

Values follow the reference's lisatools constants
(LISAanalysistools/lisatools/utils/constants.py:1-9), which are the same numbers
``few.utils.constants`` exports to the scripts (emri_pe.py:62).
"""
MSUN_SI = 1.98848e30
YRSID_SI = 31558149.763545603
MTSUN_SI = 4.925491025873693e-06
MRSUN_SI = 1476.6250615036158
PC_SI = 3.0856775814913674e16
Gpc = PC_SI * 1.0e9
PI = 3.141592653589793238462643383279502884
