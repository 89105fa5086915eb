"""Formula mean functions without ``patsy``: design matrices H(x) for the analytic-mean GP.

The CPU reference builds its design matrix with ``patsy.dmatrix(mean, data={"x": inputs.T})``
(GaussianProcess.get_design_matrix, GaussianProcess.py:485-514): the formula is written over ``x[0]``, ``x[1]``, ...
(one entry per input dimension), an intercept column comes first unless the formula removes it, an optional left-hand
side (``"y ~ ..."``) is ignored.  ``patsy`` cannot be installed offline, so this module evaluates the subset of its
formula language that makes sense for numeric inputs, with patsy's column conventions:

    terms          separated by ``+`` at bracket depth 0; ``- term`` removes a term
    intercept      present by default (first column); ``-1`` / ``+0`` / a bare ``0`` remove it, ``+1`` / ``-0`` keep it
    interaction    ``a:b`` (product of the factor columns); ``a*b`` is short for ``a + b + a:b``
    factor         any Python expression over ``x[i]`` and ``np`` -- ``x[2]``, ``I(x[0]**2)``, ``np.sin(x[1])``,
                   ``np.log(x[0] + 1.0)`` -- giving one numeric column; ``I(...)`` is the identity (it only protects
                   arithmetic operators from the formula parser, as in patsy)
    column order   intercept, then terms by number of factors, ties in order of appearance (patsy's ordering)

Not supported (``ValueError("Provided mean function is invalid")``, the reference's error): categorical factors,
formula-level parentheses / ``**`` / ``/`` / ``%in%``, multi-column factors.

``input_deriv`` returns dH/dx for the predictive derivatives (the reference GPU class adds the mean function's input
derivative, densegp_gpu.hpp:411-448 / meanfunc.hpp ``mean_inputderiv``): complex-step differentiation of the factor
expressions (exact to rounding for analytic expressions), checked against central differences and replaced by them where
an expression is not complex-analytic (``abs``, ``np.maximum`` ...).
"""
import numpy as np

_INVALID = "Provided mean function is invalid"


def _identity(v):
    return v


def _split_top(s, seps):
    """Split ``s`` on the single-character separators ``seps`` at bracket depth 0 -> [(separator before piece, piece)].
    ``**`` never splits (it is rejected at depth 0 by the caller)."""
    out, depth, start, prev = [], 0, 0, ""
    i = 0
    while i < len(s):
        c = s[i]
        if c in "([":
            if c == "(" and depth == 0:
                j = i - 1
                while j >= 0 and s[j] == " ":
                    j -= 1
                if j < 0 or not (s[j].isalnum() or s[j] in "_.])"):
                    raise ValueError(_INVALID + ": formula-level parentheses are not supported (wrap arithmetic in I(...))")
            depth += 1
        elif c in ")]":
            depth -= 1
            if depth < 0:
                raise ValueError(_INVALID + ": unbalanced brackets")
        elif depth == 0 and c in seps:
            if c == "*" and i + 1 < len(s) and s[i + 1] == "*":
                raise ValueError(_INVALID + ": '**' at formula level is not supported (write I(x[0]**2))")
            out.append((prev, s[start:i]))
            prev, start = c, i + 1
        i += 1
    if depth != 0:
        raise ValueError(_INVALID + ": unbalanced brackets")
    out.append((prev, s[start:]))
    return out


class MeanFormula(object):
    """A parsed mean-function formula: ``intercept`` (bool) and ``terms`` (tuples of factor expressions)."""

    def __init__(self, formula):
        if not isinstance(formula, str):
            raise ValueError(_INVALID)
        self.formula = formula
        rhs = formula.split("~")
        if len(rhs) > 2:
            raise ValueError(_INVALID)
        rhs = rhs[-1].strip()
        if rhs == "":
            raise ValueError(_INVALID)
        if not self._only_inside_brackets(rhs, "/") or "%in%" in rhs:
            raise ValueError(_INVALID + ": '/' and '%in%' at formula level are not supported")
        self.intercept = True
        terms = []
        for sign, piece in _split_top(rhs, "+-"):
            piece = piece.strip()
            if piece == "":
                if sign == "" and len(rhs) > 0:      # leading sign: "-1", "+0", "- x[0]"
                    continue
                raise ValueError(_INVALID)
            remove = sign == "-"
            if piece in ("0", "1"):
                keep = (piece == "1") != remove      # "+1" / "-0" keep the intercept, "+0" / "-1" remove it
                self.intercept = keep
                continue
            for term in self._expand_star(piece):
                if remove:
                    terms = [t for t in terms if frozenset(t) != frozenset(term)]
                elif all(frozenset(t) != frozenset(term) for t in terms):
                    terms.append(term)
        self.terms = sorted(terms, key=len)          # stable: ties keep their order of appearance
        self._code = {}
        for term in self.terms:
            for f in term:
                if f not in self._code:
                    try:
                        self._code[f] = compile(f, "<mean formula>", "eval")
                    except SyntaxError:
                        raise ValueError(_INVALID + ": cannot parse factor '%s'" % f)

    @staticmethod
    def _only_inside_brackets(s, ch):
        depth = 0
        for c in s:
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            elif c == ch and depth == 0:
                return False
        return True

    @staticmethod
    def _expand_star(piece):
        """``a*b:c*d`` -> list of terms (tuples of factors); ``*`` binds looser than ``:`` (patsy precedence)."""
        groups = [[tuple(dict.fromkeys(f.strip() for _, f in _split_top(g, ":")))] for _, g in _split_top(piece, "*")]
        for g in groups:
            if any(f == "" for f in g[0]):
                raise ValueError(_INVALID)
        terms = groups[0]
        for g in groups[1:]:
            crossed = []
            for a in terms:
                for b in g:
                    merged = a + tuple(f for f in b if f not in a)
                    crossed.append(merged)
            terms = terms + g + crossed
        out = []
        for t in terms:
            if all(frozenset(t) != frozenset(o) for o in out):
                out.append(t)
        return out

    # -- evaluation ----------------------------------------------------------------------------------
    @property
    def n_mean(self):
        return int(self.intercept) + len(self.terms)

    def _factor_columns(self, xT):
        """Evaluate every distinct factor over the rows of ``xT`` (D, m) -> {expression: (m,) column}."""
        env = {"np": np, "numpy": np, "x": xT, "I": _identity, "exp": np.exp, "log": np.log, "sqrt": np.sqrt,
               "sin": np.sin, "cos": np.cos, "tan": np.tan, "tanh": np.tanh, "abs": np.abs, "__builtins__": {}}
        m = xT.shape[1]
        cols = {}
        for f, code in self._code.items():
            try:
                v = eval(code, env)          # noqa: S307 -- a formula is code by design (patsy evaluates it the same way)
            except ValueError:
                raise
            except Exception as exc:       # IndexError for x[6] with D = 3, NameError, ...
                raise ValueError(_INVALID + ": factor '%s' failed (%s: %s)" % (f, type(exc).__name__, exc))
            v = np.asarray(v)
            if v.ndim == 0:
                v = np.full(m, v[()])
            if v.shape != (m,) or not (np.issubdtype(v.dtype, np.floating) or np.issubdtype(v.dtype, np.integer)
                                       or np.issubdtype(v.dtype, np.complexfloating)):
                raise ValueError(_INVALID + ": factor '%s' does not give one numeric column" % f)
            cols[f] = v
        return cols

    def _assemble(self, cols, m, dtype):
        H = np.empty((m, self.n_mean), dtype=dtype)
        k = 0
        if self.intercept:
            H[:, 0] = 1.0
            k = 1
        for term in self.terms:
            col = cols[term[0]]
            for f in term[1:]:
                col = col * cols[f]
            H[:, k] = col
            k += 1
        return H

    def design_matrix(self, inputs):
        """H (m, n_mean) for inputs (m, D)."""
        inputs = np.atleast_2d(np.asarray(inputs, dtype=np.float64))
        m = inputs.shape[0]
        xT = np.ascontiguousarray(inputs.T)
        H = self._assemble(self._factor_columns(xT), m, np.float64)
        if not np.all(np.isfinite(H)):
            raise ValueError(_INVALID + ": the design matrix is not finite")
        return H

    def _deriv_central(self, inputs, q, h):
        step = np.zeros(inputs.shape[1])
        step[q] = h
        return (self.design_matrix(inputs + step) - self.design_matrix(inputs - step)) / (2.0 * h)

    def input_deriv(self, inputs):
        """dH/dx, shape (m, D, n_mean)."""
        inputs = np.atleast_2d(np.asarray(inputs, dtype=np.float64))
        m, D = inputs.shape
        out = np.zeros((m, D, self.n_mean))
        if not self.terms:
            return out
        hc, hf = 1.0e-30, 1.0e-6
        probe = np.unique(np.linspace(0, m - 1, min(m, 8)).astype(int))
        for q in range(D):
            ok = False
            try:
                xT = np.ascontiguousarray(inputs.T).astype(np.complex128)
                xT[q] += 1j * hc
                Hc = self._assemble(self._factor_columns(xT), m, np.complex128)
                dq = Hc.imag / hc
                fd = self._deriv_central(inputs[probe], q, hf)
                ok = np.all(np.isfinite(dq)) and np.allclose(dq[probe], fd, rtol=1.0e-5, atol=1.0e-6 * (1.0 + np.abs(fd).max()))
            except (ValueError, TypeError):
                ok = False
            out[:, q, :] = dq if ok else self._deriv_central(inputs, q, hf)
        return out

    def __str__(self):
        return self.formula

    def __reduce__(self):          # compiled factor expressions are not picklable: rebuild from the formula string
        return (MeanFormula, (self.formula,))

    def __repr__(self):
        return "MeanFormula(%r)" % self.formula

    def __eq__(self, other):
        if isinstance(other, MeanFormula):
            return self.intercept == other.intercept and self.terms == other.terms
        if isinstance(other, str):
            return self.formula == other
        return NotImplemented

    def __hash__(self):
        return hash((self.intercept, tuple(self.terms)))
