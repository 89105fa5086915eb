"""Import shim for the *reference* package (TEST INFRASTRUCTURE ONLY - never imported by the product).

The reference tree at /root/reference is read-only and lacks two things needed to
`import mogp_emulator` in this container:
  * mogp_emulator/version.py  (written by the reference's setup.py at install time)
  * the third-party `patsy` package (only used for formula mean functions, which are
    out of scope here: every config uses mean=None)
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
        patsy.dmatrix = _unavailable
        patsy.dmatrices = _unavailable
        sys.modules["patsy"] = patsy
    if "mogp_emulator.version" not in sys.modules:
        ver = types.ModuleType("mogp_emulator.version")
        ver.version = "0.7.2"
        sys.modules["mogp_emulator.version"] = ver
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mogp_emulator  # noqa: E402

    return mogp_emulator
