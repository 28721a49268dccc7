"""`astropy` stand-in (TEST INFRASTRUCTURE): the hot path (`pop_cosmo_model`) never calls astropy;
`intensity_models.py` only imports `Planck18` and `astropy.units` at module level for `pop_model`."""
