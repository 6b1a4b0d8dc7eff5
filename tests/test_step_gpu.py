"""-m gpu parity of the whole per-minibatch ELBO (value of every term, gradient of every parameter) against the
oracle restatement, with identical base noise on both sides; plus training-loop behaviour."""
import copy

import numpy as np
import pytest
import torch

from oracle import elbo as oelbo

pytestmark = pytest.mark.gpu
DEV = "cuda"


from _helpers import _oracle_inputs, _setup, _small_dataset  # noqa: E402,F401


def _grad_of(p):
    """Gradient of a parameter: kernels that own a gradient (GEMM epilogues, the latent / simplex backward) write it
    straight into the optimizer's flat buffer (``_cb_grad_view``) and leave ``.grad`` unset."""
    if getattr(p, "_cb_grad_direct", False):
        return p._cb_grad_view.detach().double().cpu().clone()
    return None if p.grad is None else p.grad.detach().double().cpu()


def _product_param_grads(model):
    out = {}
    for pre, mod in (("enc_z.", model.encoder["z"]), ("enc_o.", model.encoder["other"]), ("dec.", model.decoder)):
        for k, v in mod.named_parameters():
            out[pre + k] = _grad_of(v)
    for k, v in model.param_store.unconstrained.items():
        out["ps." + k] = _grad_of(v)
    return out


@pytest.mark.parametrize("training", [True, False])
def test_elbo_terms_and_gradients_match_oracle(training, fp32_gemm):
    """fp32 library GEMMs substituted from the test side (the reference's arithmetic) + the hand-written kernels: 1e-5 on
    every term."""
    _elbo_parity(training, rtol_terms=1e-5, rtol_grads=2e-3)


@pytest.mark.parametrize("training", [True, False])
def test_elbo_tf32_tolerance(training):
    """Same comparison with the product as shipped (tcgen05 TF32 GEMMs): stated tolerance 2e-3 relative on the ELBO terms
    (TF32 keeps 10 mantissa bits in the large contractions), 3e-2 of scale on gradients."""
    _elbo_parity(training, rtol_terms=2e-3, rtol_grads=3e-2)


@pytest.mark.parametrize("model_type", ["ambient", "swapping"])
def test_elbo_other_model_types_match_oracle(model_type, fp32_gemm):
    """The rate mixtures of calculate_mu / calculate_lambda for --model ambient and swapping (model.py:25-85): same
    term-by-term and gradient-by-gradient comparison as for the default 'full' model."""
    _elbo_parity(True, rtol_terms=1e-5, rtol_grads=2e-3, model_type=model_type)


@pytest.mark.parametrize("training", [True, False])
def test_elbo_with_the_exact_convolution_likelihood_matches_oracle(training, fp32_gemm):
    """consts.USE_EXACT_LOG_PROB (model.py:336-357): the fused observation kernel with NegativeBinomialPoissonConv(
    max_poisson = 100) as the likelihood -- one-term closed form where no count is stored, the term-by-term float64 sum on
    the stored counts -- against the oracle's float64 evaluation of the same sum (without the reference's float32 cast of
    its count tables, oracle/nbpc.py), every term and every gradient."""
    _elbo_parity(training, rtol_terms=1e-5, rtol_grads=2e-3, likelihood="exact", shape=dict(exact=True))


@pytest.mark.parametrize("tf32", [False, True])
def test_elbo_matches_oracle_at_a_medium_shape(tf32, request):
    """The same term-by-term / gradient-by-gradient comparison at G = 2051 genes (not a multiple of 4: padded rows and weight
    blocks), 128 hidden units, z_dim 16 (the float4 path of the latent kernels) -- split-K products, several column tiles
    per GEMM, multi-word count bitmaps.  The oracle evaluates the reference's likelihood formula in float64 here (the
    multiprecision evaluation is per element in Python: minutes at this size)."""
    if not tf32:
        request.getfixturevalue("fp32_gemm")
    _elbo_parity(True, rtol_terms=2e-3 if tf32 else 1e-5, rtol_grads=3e-2 if tf32 else 2e-3, likelihood="reference",
                 shape=dict(n_genes=2051, z_dim=16, hidden=128), min_tensors=28)


def _elbo_parity(training, rtol_terms, rtol_grads, model_type="full", likelihood="mp", shape=None, min_tensors=28):
    ds, args, model, opt, svi, train_loader, _ = _setup(model_type, **(shape or {}))
    batch = next(iter(train_loader))
    model.train()
    with torch.no_grad():
        svi.loss_terms(batch)  # first forward fixes the encoder's offset heuristics (encoder.py:300-309)
    offset = dict(model.encoder["other"].offset)
    model.train(training)
    B, Z = batch.size(0), model.z_dim
    g = torch.Generator().manual_seed(5)
    keep = (torch.bernoulli(torch.full((B,), 0.5), generator=g) * 2.0) if training else None

    # concentrations at which gamma / beta base samples are drawn (needs the encoder's epsilon)
    inp = batch.encoder_inputs(svi._ambient_unit_vector())
    with torch.no_grad():
        enc0 = model.encoder(x=inp, chi_ambient=model.param_store["chi_ambient"].detach(), cell_prior_log=model.d_cell_loc_prior,
                             row_keep_scale=None if keep is None else keep.to(DEV))
        prob0 = enc0["p_y"].sigmoid()
        ps = model.param_store
        conc = {"phi": (ps["phi_loc"] ** 2 / ps["phi_scale"] ** 2).cpu(),
                "epsilon": ((prob0 * enc0["epsilon"] + 1 - prob0) * model.epsilon_prior).cpu(),
                "rho_alpha": (ps["rho_alpha"] if model.include_rho else torch.tensor(1.5)).cpu(),
                "rho_beta": (ps["rho_beta"] if model.include_rho else torch.tensor(50.0)).cpu()}
    noise = oelbo.draw_noise(B, Z, conc, generator=g)

    opt.zero_grad()
    chi_ambient, _ = model.ambient(B)
    terms = model.elbo_terms(batch.loader.csr, batch.row_ids, inp, noise={k: v.to(DEV) for k, v in noise.items()},
                             row_keep_scale=None if keep is None else keep.to(DEV), chi_ambient=chi_ambient)
    terms["loss"].backward()
    model.check_nan()
    got = _product_param_grads(model)

    params, buffers, cfg = _oracle_inputs(model, offset)
    x = batch.dense().double().cpu()
    ref = oelbo.elbo(x, params, buffers, cfg, {k: v.double() for k, v in noise.items()},
                     None if keep is None else keep.double(), training=training, likelihood=likelihood, model_type=model_type)
    ref["loss"].backward()

    names = [k for k in ref if not k.startswith("_")]
    for k in names:
        a, b = float(terms[k].detach()), float(ref[k].detach())
        # the small KL-type terms are sums of large cancelling fp32 pieces (e.g. 50 log 50 per droplet in
        # Gamma.log_prob) evaluated with the very torch.distributions calls the reference makes, so their fp32
        # rounding (~1e-3 absolute) is the reference's own; the observation term and the total get 1e-5 relative.
        tol = rtol_terms * max(abs(b), 1.0) + (0.0 if k in ("obs", "loss") else 3e-3)
        assert abs(a - b) <= tol, (k, a, b)
    assert abs(float(terms["loss"].detach()) - float(ref["loss"].detach())) <= rtol_terms * abs(float(ref["loss"].detach()))
    worst = {}
    for k, p in params.items():
        if p.grad is None:
            assert got[k] is None or float(got[k].abs().max()) == 0.0, k
            continue
        assert got[k] is not None, f"product has no gradient for {k}"
        scale = float(p.grad.abs().max())
        err = float((got[k] - p.grad).abs().max())
        worst[k] = err / max(scale, 1e-6)
        assert err <= rtol_grads * scale + 2e-4, (k, err, scale)  # atol: biases feeding a BatchNorm have zero true gradient
    assert len(worst) >= min_tensors  # every encoder / decoder / param-store tensor was compared


def test_training_loop_scheduler_and_graph_replay():
    """A few epochs on the small dataset: finite losses, LR follows OneCycle, parameters move, no NaN; the
    extra-step behaviour of the scheduler is the reference's (tests/test_train.py:87-108)."""
    from cellbender_b200 import train

    ds, args, model, opt, svi, train_loader, test_loader = _setup(epochs=4)
    assert svi.use_cuda_graph
    n_batches = len(train_loader)
    assert opt.total_steps == n_batches * 4
    assert abs(opt.get_last_lr()[0] - args.learning_rate * 10 / 25) < 1e-12  # starts at max_lr / 25
    p0 = opt.flat_p.clone()
    elbo_train, elbo_test = train.run_training(model, svi, train_loader, test_loader, epochs=4, test_freq=2)
    assert len(elbo_train) == 4 and len(elbo_test) == 2
    assert all(np.isfinite(elbo_train)) and all(np.isfinite(elbo_test))
    assert len(svi._graphs) >= 1  # steps after the eager warm-up were CUDA-graph replays
    assert float((opt.flat_p - p0).abs().max()) > 0
    assert opt.host_steps == n_batches * 4 and int(opt.step_count.item()) == opt.host_steps
    assert 0.9 * args.learning_rate * 10 < max(opt._lrs) <= args.learning_rate * 10 + 1e-12  # OneCycle peak = 10 lr
    assert abs(opt._lrs[-1] - args.learning_rate * 10 / 25 / 1e4) < 1e-12  # ends at initial_lr / 1e4
    with pytest.raises(ValueError, match=r"Tried to step .* times. The specified number of total steps is .*"):
        svi.step(next(iter(train_loader)))  # one step beyond total_steps raises, like torch's OneCycleLR


def test_loader_matches_reference_golden(golden_dir):
    """Device loader vs the reference DataLoader's own output for the same numpy seed (tests/golden/dataloader.npz)."""
    import os

    import scipy.sparse as sp

    from cellbender_b200.dataprep import DataLoader

    g = np.load(os.path.join(golden_dir, "dataloader.npz"))
    dataset, empties = sp.csr_matrix(g["dataset"]), sp.csr_matrix(g["empties"])
    np.random.seed(1234)
    loader = DataLoader(dataset, empties, batch_size=16, fraction_empties=0.2, shuffle=True, device=DEV)
    assert len(loader) == int(g["n_batches"])
    for epoch in (0, 1):
        batches = [b.dense().cpu().numpy() for b in loader]
        assert len(batches) == int(g["n_batches"])
        for i, b in enumerate(batches):
            np.testing.assert_array_equal(b, g[f"e{epoch}_b{i}"])


def test_training_improves_elbo_graph_and_eager_agree_statistically():
    """Same data, same seed, eager vs CUDA-graph execution: both improve the ELBO over 40 epochs at a higher
    learning rate, and end within noise of each other (the two differ only in RNG stream order)."""
    from cellbender_b200 import run, train

    finals = {}
    for use_graph in (False, True):
        ds = _small_dataset()
        args = run.default_args(epochs=40, z_dim=8, z_hidden_dims=[32], learning_rate=1e-3, use_cuda_graph=use_graph)
        model, opt, svi, tl, te = run.setup_inference(ds, args, device=DEV)
        tr, _ = train.run_training(model, svi, tl, te, epochs=40, test_freq=1000)
        first, last = float(np.mean(tr[:5])), float(np.mean(tr[-5:]))
        assert last > first + 10.0, (use_graph, first, last)
        finals[use_graph] = last
    assert abs(finals[True] - finals[False]) < 0.15 * abs(finals[False])


@pytest.mark.parametrize("use_graph", [False, True])
def test_early_weight_update_equals_separate_adam(use_graph):
    """Single-GPU training sweeps the ClippedAdam update of each large weight matrix as soon as its gradient GEMM has
    finished, overlapping the rest of backward (optim.early_block_update).  Same seed, same data: parameters and both moment buffers
    must equal those of the plain path (every gradient to HBM, one Adam sweep over the whole flat buffer at the end).  Compared after the first
    steps: biases that feed a BatchNorm have a mathematically zero gradient, Adam normalises their rounding noise to
    O(lr) steps, so later on the two runs drift apart in exactly those (irrelevant) parameters -- as any two runs do."""
    from cellbender_b200 import gemm, run, synth

    saved_min = gemm.FUSED_MIN_NUMEL
    gemm.FUSED_MIN_NUMEL = 1 << 16  # the test model's 512 x 256 layers must take the fused path too
    try:
        sim = synth.simulate_dataset(n_genes=256, cells_of_each_type=[40, 40], n_droplets=900, cell_mean_umi=2000.0,
                                     empty_mean_umi=60.0, dirichlet_alpha=0.5, seed=0, device="cpu", cell_chunk=64)
        states = {}
        n_steps = 5 if use_graph else 2  # graphs start after 3 eager warm-up steps
        for fuse in (True, False):
            ds = synth.prepare_dataset(sim, expected_cells=80, total_droplets_included=300)
            args = run.default_args(epochs=5, z_dim=16, z_hidden_dims=[512], learning_rate=1e-3, use_cuda_graph=use_graph)
            model, opt, svi, tl, _ = run.setup_inference(ds, args, device=DEV)
            opt.early_weight_updates = fuse
            blocks = [p for p in opt.params if getattr(p, "_cb_block", None) is not None]
            assert len(blocks) >= 4  # the G x 512 matrices (and the 512 x 516 / 512 x 512 hidden layers) qualify
            model.train()
            snaps = []

            def batches():
                while True:  # the loader raises StopIteration at the end of every epoch and reshuffles
                    yield from tl

            it = batches()
            for _ in range(n_steps):
                svi.step(next(it))
                torch.cuda.synchronize()
                snaps.append((opt.flat_p.clone(), opt.flat_m.clone(), opt.flat_v.clone()))
            assert int(opt.step_count.item()) == n_steps == opt.host_steps
            assert (len(svi._graphs) >= 1) == use_graph
            states[fuse] = snaps
    finally:
        gemm.FUSED_MIN_NUMEL = saved_min
    # first two steps: identical up to the last bit of the Adam arithmetic
    for k in range(2):
        for a, b, name in zip(states[True][k], states[False][k], ("param", "exp_avg", "exp_avg_sq")):
            err = float((a - b).abs().max())
            assert err <= 1e-6 * max(1.0, float(b.abs().max())), (k, name, err)
    # later steps (graph replays included): the weight matrices themselves stay together
    pa, pb = states[True][-1][0], states[False][-1][0]
    for p_, o in zip(opt.params, opt.offsets):
        blk = getattr(p_, "_cb_block", None)
        if blk is not None:
            n = blk[1] * blk[2]
            assert float((pa[o:o + n] - pb[o:o + n]).abs().max()) < 2e-3, tuple(p_.shape)


def test_loss_reader_returns_every_loss_one_step_late():
    """train.LossReader: push() hands back the previous step's loss (read from a pinned slot whose copy has finished),
    flush() the last one -- the same values loss.item() gives, without a device synchronisation per step."""
    from cellbender_b200.train import LossReader

    ds, args, model, opt, svi, train_loader, _ = _setup()
    model.train()
    reader = LossReader()
    want, got = [], []
    for epoch in range(2):
        for b in train_loader:
            loss = svi.step(b)
            want.append(loss.clone())
            v = reader.push(loss)
            if v is not None:
                got.append(v)
    got.append(reader.flush())
    assert reader.flush() is None
    assert len(got) == len(want) >= 4
    assert got == [float(w.item()) for w in want]


@pytest.mark.parametrize("use_graph", [False, True])
def test_restart_from_a_snapshot_is_bit_exact(use_graph):
    """The reference tests that training resumed from a checkpoint continues bit for bit (tests/test_checkpoint.py:353-507:
    model, optimizer, loader state and the random number generators are restored).  Same property here, in memory: nothing
    that feeds the parameters depends on atomics order (fixed-order split-K, column and block reductions), so restoring
    parameters + BatchNorm buffers + optimizer moments / step counter + loader state + numpy / torch generator states and
    training on gives exactly the parameters of the uninterrupted run -- eagerly and through the replayed step graphs."""
    import copy

    from cellbender_b200 import run, synth

    sim = synth.simulate_dataset(n_genes=256, cells_of_each_type=[40, 40], n_droplets=900, cell_mean_umi=2000.0,
                                 empty_mean_umi=60.0, dirichlet_alpha=0.5, seed=0, device="cpu", cell_chunk=64)
    ds = synth.prepare_dataset(sim, expected_cells=80, total_droplets_included=300)
    args = run.default_args(epochs=6, z_dim=16, z_hidden_dims=[64], learning_rate=1e-3, use_cuda_graph=use_graph)
    model, opt, svi, tl, _ = run.setup_inference(ds, args, device=DEV)
    model.train()

    def epochs(n):
        for _ in range(n):
            for b in tl:
                svi.step(b)
        torch.cuda.synchronize()

    epochs(2)  # past the eager warm-up steps: the graphs (if any) are captured
    snap = {"model": copy.deepcopy(model.state_dict()), "flat_p": opt.flat_p.clone(), "opt": copy.deepcopy(opt.state_dict()),
            "loader": copy.deepcopy(tl.get_state()), "served": tl._served, "np": np.random.get_state(),
            "torch": torch.get_rng_state(), "cuda": torch.cuda.get_rng_state(DEV),
            "offset": dict(model.encoder["other"].offset)}
    epochs(2)
    want = opt.flat_p.clone()
    want_m = opt.flat_m.clone()
    # restore and run the same two epochs again
    with torch.no_grad():
        opt.flat_p.copy_(snap["flat_p"])
    model.load_state_dict(snap["model"])
    opt.load_state_dict(snap["opt"])
    tl.set_state(snap["loader"]["ind_list"].copy(), snap["loader"]["ptr"])
    tl._served = snap["served"]
    np.random.set_state(snap["np"])
    torch.set_rng_state(snap["torch"])
    torch.cuda.set_rng_state(snap["cuda"], DEV)
    model.encoder["other"].offset = snap["offset"]
    epochs(2)
    assert torch.equal(opt.flat_p, want), float((opt.flat_p - want).abs().max())
    assert torch.equal(opt.flat_m, want_m)
    assert float((want - snap["flat_p"]).abs().max()) > 0


def test_evaluate_loss_replays_a_graph_and_agrees_with_eager():
    """SVI.evaluate_loss (the test pass): captured once per (loader, minibatch size) and replayed; its Monte-Carlo mean
    over repeated evaluations of one minibatch agrees with the eager evaluation's, and nothing is trained by it."""
    ds, args, model, opt, svi, train_loader, test_loader = _setup(epochs=4)
    model.train()
    for b in train_loader:
        svi.step(b)  # sets the encoder offsets, warms up
    model.eval()
    batch = next(iter(test_loader))
    p0 = opt.flat_p.clone()
    n_graphs = len(svi._graphs)
    n_eval = 300
    replayed = torch.stack([svi.evaluate_loss(batch) for _ in range(n_eval)]).double()
    assert len(svi._graphs) == n_graphs + 1
    svi.use_cuda_graph = False
    eager = torch.stack([svi.evaluate_loss(batch) for _ in range(n_eval)]).double()
    assert torch.equal(opt.flat_p, p0)
    assert float(replayed.std()) > 0  # fresh Monte-Carlo draws in every replay
    # the loss of one minibatch is a noisy Monte-Carlo estimate (one global phi draw per evaluation): the two means must
    # agree within their standard errors
    a, b_ = float(replayed.mean()), float(eager.mean())
    se = float((replayed.var() / n_eval + eager.var() / n_eval).sqrt())
    assert abs(a - b_) <= 4.0 * se + 1e-3 * abs(b_), (a, b_, se)
    assert 0.5 < float(replayed.std() / eager.std()) < 2.0
