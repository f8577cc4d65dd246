"""qldpc_b200 - B200-native Monte-Carlo decoding hot path of michelebanfi/qLDPC-branched-off.

Sub-packages mirror the reference's ``src`` tree (``codes``, ``noise``, ``decoding``,
``simulation``, ``utils``) so that ``sys.modules['src'] = qldpc_b200`` makes the reference's
``main.py`` run on the GPU backend.  All per-shot work happens in ``libqldpc_b200.so``
(hand-written sm_100a CUDA behind a C ABI, see ``include/qldpc_b200.h``); there is no CPU
fallback: importing a compute entry point without the built library raises.
"""
__version__ = "0.1.0"

_SRC_MODULES = ("codes", "codes.bb_code", "noise", "noise.builder", "noise.compiled", "noise.simulation", "noise.model",
                "noise.kernels", "noise.constants",
                "decoding", "decoding.sparse", "decoding.dense", "decoding.osd", "decoding.kernels",
                "simulation", "simulation.engine", "utils", "utils.caching", "utils.plotting", "decoding.alpha", "decoding.scopt")


def install_as_src():
    """Alias this package as the reference's ``src`` package, so that an unmodified ``main.py`` of the
    reference (``from src.simulation.engine import run_simulation`` ...) runs on the GPU backend."""
    import importlib
    import sys
    me = sys.modules[__name__]
    sys.modules["src"] = me
    for name in _SRC_MODULES:
        sys.modules["src." + name] = importlib.import_module(__name__ + "." + name)
    return me
