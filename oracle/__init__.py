"""ORACLE — test infrastructure only.

CPU restatement (numpy / torch-CPU / plain C) of the hand-object fitting hot path
of hassony2/homan (`homan/jointopt.py`, `homan/homan.py` and the losses they call)
plus the four un-vendored third-party packages the reference imports
(`neural_renderer`, `sdf`, `mano`, `libyana`).

Nothing under this directory is ever imported by the product package
`homan_b200`; only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs
of `bench.py` may use it, and only as the checker / the thing timed as the CPU
baseline.

PARITY STATUS: the in-tree reference logic (losses, gradient routing, Adam groups)
is pinned by golden vectors generated with the UNMODIFIED reference Python
(`scripts/make_golden.py`, committed under `tests/golden/`).  The third-party
kernels (NMR rasteriser, SDF grid, MANO LBS) are absent from the reference tree
and ship no tests or vectors: for those semantics parity is
UNPINNED.
"""
