"""`import finch` -- the name of the reference's Python module (lib/src/python.rs, built with pyo3 / maturin), served by
the B200 engine.  Put `<repo>/python` (and `<repo>`) on sys.path, or install the two directories side by side:

    from finch import sketch_file, Multisketch, Sketch
    a = sketch_file("a.fastq", n_hashes=1000, filter=True)
    db = Multisketch.open("refs.bsk")
    ix, best = db.best_match(a)

Everything lives in finch_rs_b200/pyfinch.py; this file only gives it the reference's name.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from finch_rs_b200.pyfinch import FinchError, Multisketch, PanicException, Sketch, SketchIter, sketch_file  # noqa: E402,F401

__all__ = ["sketch_file", "Sketch", "Multisketch", "FinchError"]
