"""Loader for the REFERENCE's own ``python/liblda/lda.py`` -- TEST INFRASTRUCTURE.

Works only where ``/root/reference`` exists (the authoring container); the GPU box
does not have it, so nothing that runs there imports this.  It is used by
``tests/golden/make_golden.py`` to generate the committed LDA fixtures and by the
CPU-only test that pins ``oracle/lda_port.py`` against the real reference.

The only incompatibility of the reference file with Python 3.12 / scipy 1.x is
``from scipy.misc import logsumexp`` (``python/liblda/lda.py:5``); we provide that
name via a shim module and load the file by path (bypassing the py2 relative
imports in ``python/liblda/__init__.py:1-2``).  The reference source is executed
where it lies; nothing is copied.
"""
import importlib.util
import os
import sys
import types

REF_LDA_PATH = "/root/reference/python/liblda/lda.py"


def available() -> bool:
    return os.path.exists(REF_LDA_PATH)


def load():
    if not available():
        raise RuntimeError("reference lda.py not present (only exists in the authoring container)")
    import scipy.special
    if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "logsumexp"):
        shim = types.ModuleType("scipy.misc")
        shim.logsumexp = scipy.special.logsumexp
        sys.modules["scipy.misc"] = shim
    spec = importlib.util.spec_from_file_location("_reference_lda", REF_LDA_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
