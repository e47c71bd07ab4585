"""One-off conversion of the reference's data tables into compact binary files shipped with the package.

  LISA_Alloc_Sh.txt (600 x 2 PSD table; FDutils.py:4)   -> data/lisa_alloc_sh.npy
  covariance.npy    (34240 x 6 MCMC samples; emri_pe.py:440) -> data/walker_covariance.npy (their 6x6 covariance)

Run in the build container (needs /root/reference):  python scripts/convert_reference_data.py
"""
import os
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "emri_frequencydomainwaveforms_b200", "data")

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    np.save(os.path.join(OUT, "lisa_alloc_sh.npy"), np.genfromtxt(os.path.join(REF, "LISA_Alloc_Sh.txt")))
    np.save(os.path.join(OUT, "walker_covariance.npy"), np.cov(np.load(os.path.join(REF, "covariance.npy")), rowvar=False))
    print("wrote", os.path.normpath(OUT))
