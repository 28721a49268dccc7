"""Distribution *descriptions* only (TEST INFRASTRUCTURE): the hot path never evaluates prior densities."""


class _D:
    def __init__(self, *args, **kw):
        self.args = args
        self.kw = kw

    def __repr__(self):
        return f"{type(self).__name__}{self.args}{self.kw}"


class TruncatedNormal(_D):
    pass


class Normal(_D):
    pass


class Uniform(_D):
    pass
