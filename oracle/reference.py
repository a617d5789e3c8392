"""Loader for the UNMODIFIED reference package installed under ``oracle/_ref`` (see oracle/build_ref.py).

TEST INFRASTRUCTURE: used by ``bench.py``'s CPU legs (``--impl reference`` and ``cpu_baseline``) and by
tests that compare against the executed reference; never by the product package.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "memento", "main.py"))


def pin_worker_threads():
    """One BLAS/OpenMP thread per worker process: the reference parallelises over genes with a process pool
    (reference main.py:397), so library threads on top of it only oversubscribe the cores (round-1 finding: the
    same run was 2.4x faster under torchrun's OMP_NUM_THREADS=1).  Worker processes inherit the environment."""
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[k] = "1"


def load():
    """Returns the reference's ``memento`` package (``memento.main`` etc. imported).  Its directory is put first on
    ``sys.path`` -- ahead of this repository's own ``memento`` alias package -- and stays there, so that joblib/loky
    workers resolve ``memento.*`` to the same files."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    if sys.path[0] != REF_DIR:
        if REF_DIR in sys.path:
            sys.path.remove(REF_DIR)
        sys.path.insert(0, REF_DIR)
    mod = sys.modules.get("memento")
    if mod is not None and not os.path.abspath(mod.__file__).startswith(REF_DIR):
        raise ImportError("another package called `memento` is already imported: %s" % mod.__file__)
    import memento
    import memento.main  # noqa: F401
    import memento.hypothesis_test  # noqa: F401
    import memento.bootstrap  # noqa: F401
    import memento.estimator  # noqa: F401
    assert os.path.abspath(memento.__file__).startswith(REF_DIR), memento.__file__
    return memento
