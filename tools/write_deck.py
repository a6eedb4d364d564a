#!/usr/bin/env python
"""Write one of the BASELINE decks as an input.nml the REFERENCE reads (cales_b200.deck.write_input), for the off-box pinning
recipe of INTEGRATION.md: run a site-built CaLES and `python -m cales_b200.run` on the same file, then tools/compare_fld.py.

  python tools/write_deck.py config2 [--nstep 100] [--ng NX NY NZ] [--dims P Q] > input.nml"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cales_b200.deck import BASELINE_DECKS, write_input  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("config", choices=sorted(BASELINE_DECKS))
    ap.add_argument("--nstep", type=int, default=100)
    ap.add_argument("--ng", type=int, nargs=3, default=None)
    ap.add_argument("--dims", type=int, nargs=2, default=(1, 1))
    a = ap.parse_args()
    d = BASELINE_DECKS[a.config]()
    if a.ng:
        d.ng = tuple(a.ng)
    d.dims = tuple(a.dims)
    d.nstep, d.isave, d.stop_type, d.is_overwrite_save = a.nstep, a.nstep, (True, False, False), True
    print(write_input(d), end="")


if __name__ == "__main__":
    main()
