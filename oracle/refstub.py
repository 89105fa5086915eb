"""Import shim for the *reference* package (TEST INFRASTRUCTURE ONLY - never imported by the product).

The reference tree at /root/reference is read-only and lacks two things needed to
`import mogp_emulator` in this container:
  * mogp_emulator/version.py  (written by the reference's setup.py at install time)
  * the third-party `patsy` package (only used for formula mean functions; the stand-in
    returns the hand-written design matrices of gp_oracle.DESIGN_FUNCTIONS for the few
    formulas the fixtures use and raises PatsyError for everything else)
This module registers in-memory stand-ins for both and puts /root/reference on sys.path.
It only exists so that tests/golden/make_golden.py can generate fixtures from the real
reference and so that the restatement in oracle/gp_oracle.py can be validated against it
*in this container*.  /root/reference does not exist on the GPU box: nothing executed
there may import this module.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MOGP_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mogp_emulator"))


def import_reference():
    """Return the imported reference package `mogp_emulator` (raises ImportError if absent)."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    if "patsy" not in sys.modules:
        patsy = types.ModuleType("patsy")

        class PatsyError(Exception):
            pass

        def _unavailable(*args, **kwargs):
            raise PatsyError("patsy is not installed (stubbed by oracle/refstub.py)")

        patsy.PatsyError = PatsyError
        patsy.ModelDesc = type("ModelDesc", (), {})
        patsy.Term = type("Term", (), {})
        patsy.EvalFactor = type("EvalFactor", (), {})
        def _dmatrix(formula, data=None, **kwargs):
            # formula mean functions of the fixtures: the design matrices patsy would build, from the oracle's hand-written
            # table (gp_oracle.DESIGN_FUNCTIONS); the reference's GP algebra then runs unmodified on top of them
            import numpy as np
            import gp_oracle
            if formula not in gp_oracle.DESIGN_FUNCTIONS or "~" in formula:
                raise PatsyError("formula %r is not in gp_oracle.DESIGN_FUNCTIONS (patsy is stubbed)" % (formula,))
            return gp_oracle.design_from_table(formula, np.asarray(data["x"]).T)

        def _dmatrices(formula, data=None, **kwargs):
            import numpy as np
            import gp_oracle
            if formula not in gp_oracle.DESIGN_FUNCTIONS:
                raise PatsyError("formula %r is not in gp_oracle.DESIGN_FUNCTIONS (patsy is stubbed)" % (formula,))
            return np.asarray(data["y"]), gp_oracle.design_from_table(formula, np.asarray(data["x"]).T)

        patsy.dmatrix = _dmatrix
        patsy.dmatrices = _dmatrices
        sys.modules["patsy"] = patsy
    if "mogp_emulator.version" not in sys.modules:
        ver = types.ModuleType("mogp_emulator.version")
        ver.version = "0.7.2"
        sys.modules["mogp_emulator.version"] = ver
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mogp_emulator  # noqa: E402

    return mogp_emulator
