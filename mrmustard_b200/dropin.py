"""Install the B200 path behind a live MrMustard (`mrmustard.math`) — see INTEGRATION.md.

The reference resolves `math.hermite_renormalized*` by name on the active backend object
(BackendManager._apply, backend_manager.py:89-116) and the jax VJP rules call `strategies.*` by name
(math/jax_vjps/hermite.py:60-74,86-102).  `install()` therefore (1) rebinds the five BackendNumpy
methods this package implements and (2) rebinds the strategy functions in
`mrmustard.math.lattice.strategies`.  `uninstall()` restores the originals.
"""
from __future__ import annotations

from . import backend, strategies

_saved: dict = {}

_STRATEGY_NAMES = ["vanilla_numba", "stable_numba", "vanilla_batch_numba", "vanilla_vjp_numba",
                   "vanilla_batch_vjp_numba", "binomial", "fast_diagonal",
                   # gate-specific strategies (BackendNumpy.displacement / beamsplitter / squeezed / squeezer look them up by name,
                   # backend_numpy.py:452-475; beamsplitter_schwinger stays the reference's numpy eigendecomposition)
                   "squeezer", "squeezed", "beamsplitter", "stable_beamsplitter", "displacement", "jacobian_displacement",
                   "grad_displacement", "beamsplitter_vjp", "squeezer_vjp", "squeezed_vjp"]


def install(fock: bool = False) -> None:
    # (`import mrmustard.math.lattice.strategies as x` fails: `mrmustard.math` is the BackendManager instance, math/__init__.py:34)
    from mrmustard.math.lattice import strategies as ref_strategies
    from mrmustard.math.backend_numpy import BackendNumpy

    if _saved:
        return
    for name in _STRATEGY_NAMES:
        _saved[("s", name)] = getattr(ref_strategies, name)
        setattr(ref_strategies, name, getattr(strategies, name))
    for name, fn in {
        "hermite_renormalized": lambda self, A, b, c, shape, stable=False, out=None:
            backend.hermite_renormalized_unbatched(A, b, c, shape, stable, out),
        "hermite_renormalized_batched": lambda self, A, b, c, shape, stable=False, out=None:
            backend.hermite_renormalized_batched(A, b, c, shape, stable, out),
        "hermite_renormalized_binomial": lambda self, A, B, C, shape, max_l2, global_cutoff:
            backend.hermite_renormalized_binomial(A, B, C, shape, max_l2, global_cutoff),
        "hermite_renormalized_diagonal": lambda self, A, B, C, cutoffs, reorderedAB:
            backend.hermite_renormalized_diagonal(A, B, C, cutoffs, reorderedAB),
        "hermite_renormalized_1leftoverMode": lambda self, A, b, c, output_cutoff, pnr_cutoffs, stable=False, reorderedAB=True:
            backend.hermite_renormalized_1leftoverMode(A, b, c, output_cutoff, pnr_cutoffs, stable, reorderedAB),
    }.items():
        _saved[("b", name)] = getattr(BackendNumpy, name)
        setattr(BackendNumpy, name, fn)
    # autoshape_numba is imported by name into the modules that call it (lab/states/base.py:383-444): rebind it wherever it lives
    import sys
    for modname in ("mrmustard.math.lattice.autoshape", "mrmustard.lab.states.base"):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "autoshape_numba"):
            _saved[("m", modname)] = mod.autoshape_numba
            mod.autoshape_numba = strategies.autoshape_numba
    if fock:   # the Fock-space contraction of `to_fock` outputs on the GPU as well
        from mrmustard.physics.ansatz.array_ansatz import ArrayAnsatz
        from . import fock as _fock

        def contract(self, other, idx1, idx2, idx_out):
            result = _fock.contract(self.array, list(idx1), other.array, list(idx2), list(idx_out))
            return ArrayAnsatz(result, batch_dims=sum(1 for label in idx_out if isinstance(label, str)))

        _saved[("a", "contract")] = ArrayAnsatz.contract
        ArrayAnsatz.contract = contract
    backend._installed = True


def uninstall() -> None:
    # (`import mrmustard.math.lattice.strategies as x` fails: `mrmustard.math` is the BackendManager instance, math/__init__.py:34)
    from mrmustard.math.lattice import strategies as ref_strategies
    from mrmustard.math.backend_numpy import BackendNumpy

    import sys
    for (kind, name), fn in _saved.items():
        if kind == "m":
            if name in sys.modules:
                sys.modules[name].autoshape_numba = fn
            continue
        if kind == "a":
            from mrmustard.physics.ansatz.array_ansatz import ArrayAnsatz
            setattr(ArrayAnsatz, name, fn)
            continue
        setattr(ref_strategies if kind == "s" else BackendNumpy, name, fn)
    _saved.clear()
    backend._installed = False
