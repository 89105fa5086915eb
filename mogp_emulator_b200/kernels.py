"""Kernel selectors accepted by the GPU emulators: ``SquaredExponential`` and ``Matern52``, by instance or
by name -- the two the reference's GPU API accepts (mogp_emulator/GaussianProcessGPU.py:268-277).  The
arithmetic lives in csrc/kmat.cu; these classes only name the family (and document its formula)."""
from . import libmogp


class _Stationary(object):
    name = None

    def __str__(self):
        return self.name + " kernel"

    __repr__ = __str__

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(self.name)


class SquaredExponential(_Stationary):
    """k(r2) = exp(-r2/2), r2 = sum_d exp(theta_d) (x_d - x'_d)^2  (Kernel.py:772-791, 946)."""
    name = "SquaredExponential"


class Matern52(_Stationary):
    """k(r2) = (1 + sqrt(5 r2) + 5/3 r2) exp(-sqrt(5 r2))  (Kernel.py:861-882, 966)."""
    name = "Matern52"


def interpret_kernel(kernel):
    """-> (libmogp.kernel_type, kernel instance); ValueError for anything else, like the reference."""
    if isinstance(kernel, str):
        name = kernel
    else:
        name = type(kernel).__name__
    if name == "SquaredExponential":
        return libmogp.kernel_type.SquaredExponential, SquaredExponential()
    if name == "Matern52":
        return libmogp.kernel_type.Matern52, Matern52()
    raise ValueError("GPU implementation requires kernel to be SquaredExponential or Matern52")
