"""TEST INFRASTRUCTURE ONLY -- import the genuine PyBMF reference as a CPU oracle.

Only `tests/`, `oracle/make_golden.py` and the dev container may use this module.
The reference lives read-only at /root/reference and does NOT exist on the GPU
box, so nothing under `-m gpu`, `smoke()` or `bench.py` may import this file.

`import PyBMF` fails in this image at PyBMF/utils/display.py:1 (matplotlib is
absent) and PyBMF/utils/evaluate_utils.py:9 (IPython is absent).  We install
inert stand-ins for the five missing plotting / notebook packages and neutralise
`show_matrix` (Asso.init_model calls it unconditionally, PyBMF/models/Asso.py:58-59).
No arithmetic on the Asso path is touched.
"""
import os
import sys
import types
import contextlib
import io

REFERENCE_ROOT = os.environ.get("PYBMF_REFERENCE_ROOT", "/root/reference")


class _Inert(types.ModuleType):
    """A module whose every attribute is a do-nothing callable / sub-module."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name == "MatplotlibDeprecationWarning":
            return DeprecationWarning

        def _noop(*a, **k):
            return None

        return _noop


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "PyBMF"))


def load():
    """Return the imported reference package (module object `PyBMF`)."""
    if "PyBMF" in sys.modules and getattr(sys.modules["PyBMF"], "_is_reference", False):
        return sys.modules["PyBMF"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colormaps",
                 "matplotlib.colors", "matplotlib.patches",
                 "IPython", "IPython.display", "p_tqdm", "mlxtend",
                 "mlxtend.frequent_patterns", "uszipcode"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Inert(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import PyBMF  # noqa
    PyBMF._is_reference = True
    import PyBMF.models.BaseModel as _bm
    import PyBMF.generators.BaseGenerator as _bg
    _bm.show_matrix = lambda *a, **k: None
    _bg.show_matrix = lambda *a, **k: None
    return PyBMF


FIT_KW = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False,
              verbose=False, display=False)


@contextlib.contextmanager
def quiet():
    """Silence the reference's print()/tqdm chatter."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield
