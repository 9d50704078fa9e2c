"""Drop-in check against the reference's own classes (runs only where /root/reference
exists, i.e. the build container): every mirrored method keeps the reference's parameter
names, order and defaults; config dataclasses keep the reference's fields and defaults."""
import dataclasses
import inspect
import os
import sys

import pytest

REF = os.environ.get("FGK_REFERENCE_SRC", "/root/reference/src")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")


@pytest.fixture(scope="module")
def mods():
    sys.path.insert(0, REF)
    import hamiltonians.molecular as rm
    import krylov.residual_expansion as rx
    import krylov.skqd as rs
    import flow_guided_krylov_b200 as f
    return rm, rx, rs, f


def _params(fn):
    # the reference's `device` default is "cuda" if torch.cuda.is_available() else "cpu": machine dependent
    return [(p.name, "<device>" if p.name == "device" else p.default)
            for p in inspect.signature(fn).parameters.values() if p.name != "self"]


def _assert_superset(ref_fn, my_fn, what):
    rp, mp = _params(ref_fn), _params(my_fn)
    assert mp[:len(rp)] == rp, f"{what}: reference {rp} vs ours {mp}"
    for name, default in mp[len(rp):]:
        assert default is not inspect.Parameter.empty, f"{what}: extra parameter {name} needs a default"


def test_hamiltonian_interface(mods):
    rm, _, _, f = mods
    for m in ["__init__", "diagonal_elements_batch", "diagonal_element", "get_connections",
              "get_all_connections_with_indices", "get_connections_parallel", "matrix_elements_fast",
              "matrix_elements", "get_sparse_matrix_elements", "get_hf_state", "fci_energy", "_config_to_index"]:
        _assert_superset(getattr(rm.MolecularHamiltonian, m), getattr(f.MolecularHamiltonian, m),
                         f"MolecularHamiltonian.{m}")
    # the hook utils/connection_cache.py:250 looks for (the reference itself does not define it)
    assert hasattr(f.MolecularHamiltonian, "get_connections_batch")
    assert [x.name for x in dataclasses.fields(rm.MolecularIntegrals)] == \
           [x.name for x in dataclasses.fields(f.MolecularIntegrals)]


def test_expander_interface(mods):
    _, rx, _, f = mods
    for cls in ("SelectedCIExpander", "ResidualBasedExpander"):
        for m in ["__init__", "expand_basis", "_diagonalize"]:
            _assert_superset(getattr(getattr(rx, cls), m), getattr(getattr(f, cls), m), f"{cls}.{m}")
    _assert_superset(rx.SelectedCIExpander._find_important_configs,
                     f.SelectedCIExpander._find_important_configs, "_find_important_configs")
    ref = {x.name: x.default for x in dataclasses.fields(rx.ResidualExpansionConfig)}
    mine = {x.name: x.default for x in dataclasses.fields(f.ResidualExpansionConfig)}
    assert ref == mine


def test_skqd_interface(mods):
    _, _, rs, f = mods
    for m in ["__init__", "generate_krylov_samples", "build_cumulative_basis", "get_basis_states",
              "compute_ground_state_energy", "run"]:
        _assert_superset(getattr(rs.SampleBasedKrylovDiagonalization, m),
                         getattr(f.SampleBasedKrylovDiagonalization, m), f"SKQD.{m}")
    for m in ["__init__", "get_combined_basis", "run_with_nf"]:
        _assert_superset(getattr(rs.FlowGuidedSKQD, m), getattr(f.FlowGuidedSKQD, m), f"FlowGuidedSKQD.{m}")
    ref = {x.name: x.default for x in dataclasses.fields(rs.SKQDConfig)}
    mine = {x.name: x.default for x in dataclasses.fields(f.SKQDConfig)}
    for k, v in ref.items():
        assert k in mine and mine[k] == v, f"SKQDConfig.{k}"


def test_integral_front_end_interface(mods):
    """compute_molecular_integrals and the molecule factories keep the reference's names,
    parameters and default geometries (molecular.py:945-1139)."""
    rm, _, _, f = mods
    _assert_superset(rm.compute_molecular_integrals, f.compute_molecular_integrals, "compute_molecular_integrals")
    for name in ("create_h2_hamiltonian", "create_lih_hamiltonian", "create_h2o_hamiltonian",
                 "create_beh2_hamiltonian", "create_nh3_hamiltonian", "create_n2_hamiltonian",
                 "create_ch4_hamiltonian"):
        _assert_superset(getattr(rm, name), getattr(f, name), name)


def test_mirror_classes_accept_the_reference_config_objects(mods):
    """pipeline.py hands its OWN ResidualExpansionConfig / SKQDConfig instances to the classes it
    instantiates (pipeline.py:505-515, :700-712): every `self.config.<field>` our mirrors read
    must exist on the reference's dataclasses."""
    import re
    _, rx, rs, f = mods
    import flow_guided_krylov_b200.expansion as fexp
    root = os.path.dirname(os.path.abspath(fexp.__file__))
    for fname, ref_cfg in (("expansion.py", rx.ResidualExpansionConfig()), ("skqd.py", rs.SKQDConfig())):
        src = open(os.path.join(root, fname)).read()
        for attr in set(re.findall(r"self\.config\.(\w+)", src)):
            assert hasattr(ref_cfg, attr), f"{fname}: self.config.{attr} is not a field of the reference config"


def test_integration_recipe_of_INTEGRATION_md(mods):
    """The no-edit integration of INTEGRATION.md section 3, as far as it goes without a GPU: the
    combined class is constructible (MRO) and passes pipeline.py:318's isinstance check at class
    level; the names the recipe patches exist where pipeline.py looks them up."""
    rm, rx, rs, f = mods

    class MolecularHamiltonian(f.MolecularHamiltonian, rm.MolecularHamiltonian):
        def __init__(self, integrals, device="cuda"):
            f.MolecularHamiltonian.__init__(self, integrals, device)

    assert issubclass(MolecularHamiltonian, rm.MolecularHamiltonian)
    # our methods win over the reference's in the MRO
    for m in ("get_connections", "diagonal_elements_batch", "matrix_elements_fast", "fci_energy", "get_hf_state"):
        assert getattr(MolecularHamiltonian, m) is getattr(f.MolecularHamiltonian, m)
    assert hasattr(rx, "SelectedCIExpander") and hasattr(rx, "ResidualBasedExpander") and hasattr(rs, "FlowGuidedSKQD")
    src = open(os.path.join(REF, "pipeline.py")).read()
    assert "from krylov.residual_expansion import SelectedCIExpander" in src      # late import: patch the module
    assert "isinstance(hamiltonian, MolecularHamiltonian)" in src
    assert "skqd = FlowGuidedSKQD(" in src                                         # module-level name: patch pipeline.FlowGuidedSKQD
