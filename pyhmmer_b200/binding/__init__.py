"""The binding into an installed pyhmmer (`pyhmmer_cuda`, a Cython extension built by `build.py` in this directory) and the
one-call form of it: `hmmsearch`, `hmmscan`, `phmmer`, `jackhmmer`, `nhmmer` below are pyhmmer's own functions
(src/pyhmmer/hmmer/__init__.py) run with `pyhmmer_cuda.install()` in force -- same arguments, same `TopHits`, the searches
on the GPU -- i.e. what SURVEY 8(b) calls ``hmmsearch(..., backend="cuda")``.

`install()` swaps the `pipeline_class` attribute of pyhmmer's worker classes (process-wide); the wrappers keep it swapped
until their iterator is exhausted or closed, then restore what was there.  Use `pyhmmer_cuda.install()` directly to keep it
for a whole program.
"""
import functools


def _on_cuda(name):
    def run(*args, **kwargs):
        import pyhmmer
        from . import pyhmmer_cuda
        undo = pyhmmer_cuda.install()
        try:
            yield from getattr(pyhmmer, name)(*args, **kwargs)
        finally:
            undo()
    run.__name__ = run.__qualname__ = name
    run.__doc__ = "`pyhmmer.%s` with the searches on the GPU (see the module docstring); returns an iterator of `TopHits`." % name
    return run


hmmsearch = _on_cuda("hmmsearch")
hmmscan = _on_cuda("hmmscan")
phmmer = _on_cuda("phmmer")
jackhmmer = _on_cuda("jackhmmer")
nhmmer = _on_cuda("nhmmer")
