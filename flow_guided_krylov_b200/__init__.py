"""flow_guided_krylov_b200 -- B200-native determinant-space Hamiltonian engine
behind the call signatures of George930502/Flow-Guided-Krylov's hot path:

    MolecularHamiltonian / MolecularIntegrals   (reference src/hamiltonians/molecular.py)
    SelectedCIExpander / ResidualBasedExpander  (reference src/krylov/residual_expansion.py)
    SampleBasedKrylovDiagonalization / FlowGuidedSKQD / SKQDConfig (reference src/krylov/skqd.py)
    compute_molecular_integrals / create_*_hamiltonian  (reference molecular.py:945-1139, without PySCF:
                                                 STO-3G + RHF in numpy, host-side set-up only)

All compute is hand-written sm_100a CUDA in csrc/ behind the C ABI of
include/fgk_b200.h.  There is no CPU fallback: importing is cheap, but any call
without the built library and a CUDA device raises.
"""
from ._native import H_DROP_ZEROS, H_FLAT_WALK, H_HASH_WALK, H_RAW, H_SYM, PT2_MAXABS, PT2_SUM  # noqa: F401
from .hamiltonian import (BasisIndex, MolecularHamiltonian, MolecularIntegrals,  # noqa: F401
                          ProjectedH, sort_unique_dets)
from .expansion import (Pt2Workspace, ResidualBasedExpander, ResidualExpansionConfig,  # noqa: F401
                        SelectedCIExpander, default_pt2_workspace, pt2_candidates, pt2_select, select_top_k)
from .skqd import FlowGuidedSKQD, SampleBasedKrylovDiagonalization, SKQDConfig  # noqa: F401
from .solvers import expm_multiply, lowest_eigenpairs  # noqa: F401
from .sto3g import (compute_molecular_integrals, create_beh2_hamiltonian, create_ch4_hamiltonian,  # noqa: F401
                    create_h2_hamiltonian, create_h2o_hamiltonian, create_lih_hamiltonian,
                    create_n2_hamiltonian, create_nh3_hamiltonian)

__version__ = "0.1.0"
