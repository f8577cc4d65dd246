"""qldpc_b200 - B200-native Monte-Carlo decoding hot path of michelebanfi/qLDPC-branched-off.

Sub-packages mirror the reference's ``src`` tree (``codes``, ``noise``, ``decoding``,
``simulation``, ``utils``) so that ``sys.modules['src'] = qldpc_b200`` makes the reference's
``main.py`` run on the GPU backend.  All per-shot work happens in ``libqldpc_b200.so``
(hand-written sm_100a CUDA behind a C ABI, see ``include/qldpc_b200.h``); there is no CPU
fallback: importing a compute entry point without the built library raises.
"""
__version__ = "0.1.0"
