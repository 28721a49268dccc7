class _Planck18:
    def __getattr__(self, name):
        raise RuntimeError("astropy is not installed; Planck18 is a placeholder (pop_cosmo_model never uses it)")


Planck18 = _Planck18()
