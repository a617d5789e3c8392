"""memento_b200: B200-native drop-in for memento's estimation + bootstrap-testing hot path.

    import memento_b200 as memento
    memento.setup_memento(adata, q_column='q')
    memento.create_groups(adata, label_columns=['stim', 'cell'])
    memento.compute_1d_moments(adata)
    memento.ht_1d_moments(adata, covariate=cov, treatment=tr, num_boot=10000, resampling='bootstrap')

Same call signatures and ``adata.uns['memento']`` schema as the reference (memento/main.py).
"""
from .main import (compute_1d_moments, compute_2d_moments, create_groups, get_corr_matrix,  # noqa: F401
                   ht_1d_moments, ht_2d_moments, setup_memento)
from .getters import (fdrcorrect, get_1d_ht_result, get_1d_moments, get_2d_ht_result, get_2d_moments,  # noqa: F401
                      get_groups, prepare_to_save)
from .anndata_lite import AnnDataLite  # noqa: F401
from .io import load_results, read_10x_mtx, save_results, write_10x_mtx  # noqa: F401
from . import simulate  # noqa: F401  (reference memento/simulate.py semantics, on the device)

__version__ = "0.1.0"
