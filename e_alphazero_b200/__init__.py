"""B200-native (sm_100a) E-MCTS self-play hot path: a drop-in for the
emctx / pgx calls of emcts/e-alphazero's selfplay.py and reanalyze.py.

The CUDA library is loaded lazily on first use (``_lib.load()``) and raises
if it is missing -- there is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
