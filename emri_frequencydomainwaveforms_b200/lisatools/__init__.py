"""GPU-backed mirrors of the lisatools entry points on the FD path (inner_product, snr, Likelihood)."""
