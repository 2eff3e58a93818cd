"""B200-native Asso hot path of PyBMF.  `pybmf_b200.models` / `pybmf_b200.utils` mirror the module
paths `PyBMF.models` / `PyBMF.utils` for the classes and helpers on the path."""
import sys as _sys


def install_as_pybmf(force: bool = False):
    """Make `import PyBMF`, `from PyBMF.models import Asso, AssoIter`, `from PyBMF.utils import matmul, TP, ...` resolve to
    this package, so that a script written against the reference runs unchanged on the B200 path:

        import pybmf_b200; pybmf_b200.install_as_pybmf()
        from PyBMF.models import Asso            # -> pybmf_b200.models.Asso

    Refuses to shadow a genuine PyBMF that is already imported unless force=True."""
    from . import models, utils
    this = _sys.modules[__name__]
    have = _sys.modules.get("PyBMF")
    if have is not None and have is not this and not force:
        raise RuntimeError("a different PyBMF package is already imported; pass force=True to shadow it")
    _sys.modules["PyBMF"] = this
    _sys.modules["PyBMF.models"] = models
    _sys.modules["PyBMF.utils"] = utils
    return this
