"""The reference-side binding (cellbender_b200/adapter.py): mechanics of the four call-site patches inside the UNMODIFIED
reference, checked in the build container where /root/reference is importable through oracle/ref_shim.py (no GPU there, so
no compute calls; the GPU box has no reference, so there the whole module is skipped -- parity of the patched-in
implementations is what the -m gpu tests establish on the recorded reference fixtures)."""
import inspect

import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="the reference is not present on this machine")


@pytest.fixture
def installed():
    ref_shim.install()
    from cellbender_b200 import adapter

    patched = adapter.install()
    yield adapter, patched
    adapter.uninstall()


def test_adapter_patches_the_four_call_sites(installed):
    adapter, patched = installed
    import cellbender.remove_background.data.dataprep as ref_dp
    import cellbender.remove_background.distributions.NegativeBinomialPoissonConvApprox as ref_nb
    import cellbender.remove_background.estimation as ref_est
    import cellbender.remove_background.posterior as ref_post

    from cellbender_b200 import estimation

    assert ref_dp.DataLoader is adapter.ResidentDataLoader
    assert ref_nb.NegativeBinomialPoissonConvApprox.log_prob is adapter._nbpc_approx_log_prob
    for cls in ("MAP", "Mean", "ThresholdCDF", "SingleSample", "MultipleChoiceKnapsack"):
        assert getattr(ref_est, cls).estimate_noise is getattr(estimation, cls).estimate_noise
    assert ref_post.Posterior._get_cell_noise_count_posterior_coo is adapter._get_cell_noise_count_posterior_coo
    assert len(patched) >= 8


def test_signatures_are_call_compatible_with_the_reference(installed):
    """Every argument the reference passes at the call site is accepted by the replacement, with the same name."""
    adapter, patched = installed
    originals = {(getattr(o, "__name__", str(o)), n): orig for o, n, orig in patched}

    def params(f):
        return list(inspect.signature(f).parameters)

    # DataLoader(dataset, empty_drop_dataset, batch_size, fraction_empties, shuffle, sort_by, use_cuda)  dataprep.py:47-60
    ref_loader = [v for (o, n), v in originals.items() if n == "DataLoader"][0]
    assert params(ref_loader.__init__) == params(adapter.ResidentDataLoader.__init__)
    # estimate_noise(noise_log_prob_coo, noise_offsets, **kwargs)  estimation.py:42-52; kwargs of run.py:316-323
    for (o, n), orig in originals.items():
        if n == "estimate_noise":
            new = params(getattr(__import__("cellbender_b200.estimation", fromlist=[o]), o).estimate_noise)
            assert new[:3] == ["self", "noise_log_prob_coo", "noise_offsets"]
            assert "kwargs" in new or set(params(orig)) <= set(new)
    # _get_cell_noise_count_posterior_coo(n_samples, y_map, n_counts_max, smallest_log_probability)  posterior.py:424-428
    orig = [v for (o, n), v in originals.items() if n == "_get_cell_noise_count_posterior_coo"][0]
    assert params(orig) == params(adapter._get_cell_noise_count_posterior_coo)
    for name, p_ref in inspect.signature(orig).parameters.items():
        assert inspect.signature(adapter._get_cell_noise_count_posterior_coo).parameters[name].default == p_ref.default


def test_uninstall_restores_the_reference():
    ref_shim.install()
    import cellbender.remove_background.estimation as ref_est

    from cellbender_b200 import adapter

    before = ref_est.MAP.__dict__["estimate_noise"]
    adapter.install()
    assert ref_est.MAP.__dict__["estimate_noise"] is not before
    adapter.uninstall()
    assert ref_est.MAP.__dict__["estimate_noise"] is before
