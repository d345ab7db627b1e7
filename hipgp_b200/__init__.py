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


class _DropInFinder:
    """Meta-path finder / loader that answers the reference's module names with this package's modules.  Going through
    the import system (instead of seeding sys.modules) makes `import ziggy.misc.toeplitz_tensor`, `from ziggy.misc import
    toeplitz_expanded` and `from ziggy import kernels` all work: the parent packages are imported from the reference tree
    as usual and get the drop-in bound as their attribute."""

    def __init__(self, table):
        self.table = table

    def find_spec(self, fullname, path=None, target=None):
        if fullname in self.table:
            import importlib.util
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return self.table[spec.name]

    def exec_module(self, module):
        pass


def install_as_ziggy():
    """Make `import ziggy.misc.toeplitz_tensor` etc. resolve to this package's drop-ins.  Call it before the reference's
    modules are imported; the `ziggy` package itself (its model classes, `ziggy.misc.util`, `ziggy.misc.stats` ...) keeps
    coming from wherever the reference lives on sys.path."""
    from . import toeplitz_tensor, toeplitz_expanded, cg, _inv_matmul, kernels
    table = {"ziggy.misc.toeplitz_tensor": toeplitz_tensor, "ziggy.misc.toeplitz_expanded": toeplitz_expanded,
             "ziggy.misc.cg": cg, "ziggy.misc._inv_matmul": _inv_matmul, "ziggy.kernels": kernels}
    for f in list(sys.meta_path):
        if isinstance(f, _DropInFinder):
            sys.meta_path.remove(f)
    sys.meta_path.insert(0, _DropInFinder(table))
    for name, mod in table.items():           # already imported from the reference?  rebind the names
        if name in sys.modules:
            sys.modules[name] = mod
            parent, _, leaf = name.rpartition(".")
            if parent in sys.modules:
                setattr(sys.modules[parent], leaf, mod)
