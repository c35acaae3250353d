"""Host-side matrix helpers of the results API (mirror of reference cpflow/matrix_utils.py:11-42).
The optimisation never calls these: losses inside the loop are evaluated by the CUDA engine."""
import numpy as np


def theoretical_lower_bound(n):
    """Minimum number of CNOT gates to decompose an arbitrary n-qubit unitary (matrix_utils.py:11-14)."""
    return int((4 ** n - 3 * n - 1) / 4 + 1)


def trace_prod(u, v):
    """Tr(U^dagger V) (matrix_utils.py:17-23)."""
    return (np.conj(u) * v).sum()


def disc(u, u_target):
    """matrix_utils.py:26-32."""
    return 1 - np.abs(trace_prod(u, u_target)) / u_target.shape[0]


def cost_HST(u, u_target):
    """1 - |Tr(U V^dagger)|^2 / N^2 (matrix_utils.py:35-42)."""
    n = u_target.shape[0]
    return 1 - np.abs((u * np.conj(u_target)).sum()) ** 2 / n ** 2
