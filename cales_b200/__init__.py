"""cales_b200: B200-native (sm_100a) implementation of the CaLES per-RK3-substep hot path.

The product is libcales_b200.so (hand-written CUDA behind the C ABI of include/cales_b200.h);
this package is its host-side mirror of the reference's solver interface.  It never imports
`oracle/` and has no CPU fallback."""
from .deck import Deck, deck_cavity, deck_channel, deck_duct, deck_tgv, read_input  # noqa: F401

__all__ = ["Deck", "deck_channel", "deck_tgv", "deck_duct", "deck_cavity", "read_input"]
