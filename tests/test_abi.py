"""CPU checks of the boundary: the C-ABI library loads and exports every symbol the header declares,
and the product refuses to run without a GPU (no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from memento_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "memento_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), name
    assert lib.mm_version() >= 100


def test_python_signatures_cover_the_header():
    names = set(declared_symbols()) - {"mm_version", "mm_last_error"}
    bound = set(_lib.SIGNATURES) | set(_lib.HOST_SIGNATURES)
    assert names == bound, names ^ bound


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "memento_b200.h"\nint main(void){return mm_version() == 0;}\n')
    rc = os.system("gcc -std=c99 -Wall -Werror -fsyntax-only -I%s %s" % (os.path.join(ROOT, "include"), src))
    assert rc == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import memento_b200 as memento
    from memento_b200 import synth
    ad = synth.make_counts(50, 20, seed=1)
    with pytest.raises(_lib.MementoCudaError):
        memento.setup_memento(ad, "q")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "scrna-parameter-estimation_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(base, f)


def test_poisson_table_offsets_host_helper():
    """Host-only entry point (no GPU needed): table k-ranges must hold all but ~2^-32 of the mass."""
    import scipy.stats as st
    off, total = _lib.poisson_table_offsets(64)
    assert off[0] == 0 and off[1] == 1 and total > off[64] > off[63]
    for n in (1, 7, 64):
        length = (off[n + 1] if n < 64 else total) - off[n]
        hi = int(np.floor(n)) + int(np.ceil(7.5 + np.sqrt(44.4 * n + 56.0)))
        lo = max(0, int(np.floor(n - np.sqrt(44.4 * n))) - 1)
        assert length == hi - lo + 1
        assert st.poisson.sf(hi, n) < 2.4e-10 and (lo == 0 or st.poisson.cdf(lo - 1, n) < 2.4e-10)
