"""Placeholder for the licensed `gauopen` package (absent): the same idea the reference's own
docs build uses (docs/source/conf.py:9-28) so that gauNEGF.matTools / surfGTester import."""


class _Placeholder:
    def __getattr__(self, _name):
        return self

    def __call__(self, *args, **kwargs):
        return self


QCOpMat = _Placeholder()
QCBinAr = _Placeholder()
QCUtil = _Placeholder()
