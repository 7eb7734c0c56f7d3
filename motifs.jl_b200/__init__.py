"""motifs.jl_b200 — B200 (sm_100a) implementation of the data-parallel hot path of kchu25/MOTIFs.jl.

The package directory is named after the reference (`motifs.jl_b200`), which is not a valid Python
identifier; import it as ``motifs_jl_b200`` through the shim at the repository root.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (include/motifs_b200.h) -> lib/libmotifs_b200.so
  _lib.py      ctypes binding of the C ABI (the same symbols a Julia `ccall` binds)
  inference.py host-side mirror of the reference's scan / filter / count / Fisher entry points
There is no CPU fallback: every entry point raises if the CUDA library is missing or no GPU exists.
"""
from . import _lib  # noqa: F401
from ._lib import Context, Sequences, MB200Error, library_path  # noqa: F401


def discover_motifs(datapath, save_path, num_epochs=None, **kw):
    """discover_motifs(datapath, save_path; num_epochs) — src/wrap.jl:1-11."""
    from .wrap import discover_motifs as _dm
    return _dm(datapath, save_path, num_epochs=num_epochs, **kw)
