"""CPU oracle for the memento estimation + bootstrap-testing hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy/scipy restatement of the reference
algorithm (atarashansky/scrna-parameter-estimation, package ``memento``), function by function,
each citing the reference file:line it follows.  It exists so that the CUDA product path can be
checked against something that runs on any host.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``scrna-parameter-estimation_b200/memento_b200``) never does.

Pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the UNMODIFIED reference executed in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference/memento`` through a stub shim and
writes ``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` replays every fixture through
this restatement.
"""
