"""hipgp_b200 -- B200-native (sm_100a) implementation of HIP-GP's structured-kernel hot path behind the `ziggy` API.

  hipgp_b200.toeplitz_tensor.ToeplitzTensor      <-> ziggy.misc.toeplitz_tensor.ToeplitzTensor
  hipgp_b200.toeplitz_expanded.ToeplitzMatmul    <-> ziggy.misc.toeplitz_expanded.ToeplitzMatmul, gram_solve
  hipgp_b200.cg.conj_grad / conj_grad2           <-> ziggy.misc.cg
  hipgp_b200._inv_matmul.InvMatmul               <-> ziggy.misc._inv_matmul.InvMatmul
  hipgp_b200.kernels                             <-> ziggy.kernels (+ k/kprime of ziggy.exact_gp_1d_derivatives)
  hipgp_b200.hipgp.ToeplitzInducingGP            <-> compute_kn / _make_grams of ziggy.hipgp / ziggy.svi_gp

`install_as_ziggy()` registers these modules under the reference's module names so that unmodified reference code
(`ziggy/hipgp.py`, the experiment scripts) imports them instead of the torch implementations.
"""
import sys

__version__ = "0.1.0"


def install_as_ziggy():
    """Make `import ziggy.misc.toeplitz_tensor` etc. resolve to this package's drop-ins (call before importing ziggy)."""
    from . import toeplitz_tensor, toeplitz_expanded, cg, _inv_matmul, kernels
    sys.modules["ziggy.misc.toeplitz_tensor"] = toeplitz_tensor
    sys.modules["ziggy.misc.toeplitz_expanded"] = toeplitz_expanded
    sys.modules["ziggy.misc.cg"] = cg
    sys.modules["ziggy.misc._inv_matmul"] = _inv_matmul
    sys.modules["ziggy.kernels"] = kernels
