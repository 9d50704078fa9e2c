"""The reference's UNCHANGED pipeline (src/pipeline.py: FlowGuidedKrylovPipeline, PipelineConfig)
run through the engine, exactly as INTEGRATION.md section 3 prescribes, next to the unmodified
reference on the CPU (BASELINE.json configs[0]: LiH STO-3G, full pipeline).

The reference is pure Python and is vendored, unmodified, into the git-ignored oracle/_ref/ by
tools/vendor_ref.sh (it travels to the GPU box with the gpurun snapshot); `normflows` is absent
from the image and replaced by the 20-line stub in tests/stubs/ (the molecular path never
instantiates its classes).  Skipped when oracle/_ref/src is not there."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "oracle", "_ref", "src")
F32_ENVELOPE = 2e-5


@pytest.fixture(scope="module")
def ref():
    if not os.path.isdir(REF_SRC):
        pytest.skip("oracle/_ref/src missing: run tools/vendor_ref.sh where /root/reference exists")
    for p in (REF_SRC, os.path.join(ROOT, "tests", "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import hamiltonians.molecular as ref_mol
    import krylov.residual_expansion as ref_exp
    import krylov.skqd as ref_skqd
    import pipeline
    return dict(mol=ref_mol, exp=ref_exp, skqd=ref_skqd, pipeline=pipeline)


def _integrals(name):
    from flow_guided_krylov_b200 import sto3g
    geo = {"lih": sto3g.lih_geometry, "beh2": sto3g.beh2_geometry}[name]()
    return sto3g.compute_molecular_integrals(geo)


def _quiet_run(p):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
        res = p.run(progress=False)
    return res, buf.getvalue()


def _patched(ref, monkeypatch):
    """INTEGRATION.md section 3, verbatim."""
    import flow_guided_krylov_b200 as fgk
    ref_mol, ref_exp, ref_skqd, pipeline = ref["mol"], ref["exp"], ref["skqd"], ref["pipeline"]

    class MolecularHamiltonian(fgk.MolecularHamiltonian, ref_mol.MolecularHamiltonian):
        def __init__(self, integrals, device="cuda"):
            fgk.MolecularHamiltonian.__init__(self, integrals, device)

    monkeypatch.setattr(ref_exp, "SelectedCIExpander", fgk.SelectedCIExpander)
    monkeypatch.setattr(ref_exp, "ResidualBasedExpander", fgk.ResidualBasedExpander)
    monkeypatch.setattr(pipeline, "FlowGuidedSKQD", fgk.FlowGuidedSKQD)
    monkeypatch.setattr(ref_skqd, "FlowGuidedSKQD", fgk.FlowGuidedSKQD)
    return MolecularHamiltonian


def test_unchanged_pipeline_lih_engine_vs_cpu_reference(ref, monkeypatch):
    """configs[0]: PipelineConfig() untouched.  LiH reaches its FCI energy after Stage 3 in both
    arms (Stage 4 is skipped by pipeline.py:657-675 for bases < 300), so the final energies must
    agree within the float32 envelope although the NF samples differ (CPU vs CUDA RNG)."""
    from oracle import oracle as orc
    integ = _integrals("lih")
    pipeline, ref_mol = ref["pipeline"], ref["mol"]
    # (1) the unmodified reference on the CPU
    torch.manual_seed(0)
    np.random.seed(0)
    ri = ref_mol.MolecularIntegrals(integ.h1e, integ.h2e, integ.nuclear_repulsion, integ.n_electrons,
                                    integ.n_orbitals, integ.n_alpha, integ.n_beta)
    H_ref = ref_mol.MolecularHamiltonian(ri, device="cpu")
    res_ref, _ = _quiet_run(pipeline.FlowGuidedKrylovPipeline(H_ref, pipeline.PipelineConfig(device="cpu")))
    # (2) the same pipeline code over the engine
    MolecularHamiltonian = _patched(ref, monkeypatch)
    torch.manual_seed(0)
    np.random.seed(0)
    H = MolecularHamiltonian(integ, device="cuda")
    assert isinstance(H, ref_mol.MolecularHamiltonian)               # pipeline.py:318
    p = pipeline.FlowGuidedKrylovPipeline(H, pipeline.PipelineConfig(device="cuda"))
    assert p.is_molecular
    res, log = _quiet_run(p)
    for key in ("nf_nqs_energy", "nf_basis_size", "residual_expansion_stats", "residual_energy",
                "skqd_energy", "combined_energy"):
        assert key in res and key in res_ref, key
    O = orc.OracleHam(integ.h1e.astype(np.float32), integ.h2e.astype(np.float32), integ.n_alpha, integ.n_beta,
                      float(integ.nuclear_repulsion))
    E_fci, _ = O.diagonalize(O.fci_basis())
    assert abs(res_ref["combined_energy"] - E_fci) < F32_ENVELOPE     # the reference reaches FCI ...
    assert abs(res["combined_energy"] - E_fci) < 1e-6                 # ... and so does the engine (FP64)
    assert abs(res["combined_energy"] - res_ref["combined_energy"]) < F32_ENVELOPE
    assert abs(res["combined_energy"] - (-7.96379759)) < F32_ENVELOPE  # published pin, SKQD_VALIDATION_REPORT.md:87
    st = res["residual_expansion_stats"]
    assert st["final_basis_size"] >= st["initial_basis_size"] and st["final_energy"] <= st["initial_energy"] + 1e-9
    assert "Stage 3" in log


def test_unchanged_pipeline_beh2_runs_stage4_on_the_engine(ref, monkeypatch):
    """BeH2 (1,225 determinants): a short Stage-1 budget leaves work for Stage 3 and Stage 4, both of
    which then run on the engine's classes with the REFERENCE's own config objects
    (ResidualExpansionConfig / SKQDConfig built inside pipeline.py)."""
    from oracle import oracle as orc
    integ = _integrals("beh2")
    pipeline = ref["pipeline"]
    MolecularHamiltonian = _patched(ref, monkeypatch)
    torch.manual_seed(1)
    np.random.seed(1)
    H = MolecularHamiltonian(integ, device="cuda")
    cfg = pipeline.PipelineConfig(device="cuda", max_epochs=120, min_epochs=40, samples_per_batch=1000,
                                  max_krylov_dim=4, shots_per_krylov=20000)
    p = pipeline.FlowGuidedKrylovPipeline(H, cfg, auto_adapt=False)
    res, log = _quiet_run(p)
    O = orc.OracleHam(integ.h1e.astype(np.float32), integ.h2e.astype(np.float32), integ.n_alpha, integ.n_beta,
                      float(integ.nuclear_repulsion))
    E_fci, _ = O.diagonalize(O.fci_basis())
    assert res["combined_energy"] >= E_fci - 1e-6                     # variational (+1e-8 regularisation)
    assert res["combined_energy"] <= res["residual_energy"] + 1e-9
    assert res["residual_energy"] <= res["residual_expansion_stats"]["initial_energy"] + 1e-9
    if not res.get("skqd_skipped"):
        sk = res["skqd_results"]
        for key in ("krylov_dims", "energies_krylov", "energies_combined", "basis_sizes_krylov",
                    "basis_sizes_combined", "energy_nf_only", "nf_basis_size", "best_stable_energy"):
            assert key in sk, key
        assert sk["best_stable_energy"] >= E_fci - 1e-6
        assert len(sk["energies_combined"]) == cfg.max_krylov_dim - 1
    assert abs(res["combined_energy"] - E_fci) < 5e-3                 # within a few mHa of FCI
