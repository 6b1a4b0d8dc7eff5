"""Posterior generation on the GPU behind the reference's ``Posterior`` interface (reference
cellbender/remove_background/posterior.py:170-980): same constructor, same ``latents_map`` /
``cell_noise_count_posterior_coo(**kwargs)`` / ``compute_denoised_counts(estimator_constructor, **kwargs)`` methods,
same return types (scipy COO keyed by m = n * G_all + g plus the offsets dict).

Differences in HOW:
  * the encoder runs once per cell chunk, not once per latent sample (it is deterministic in eval mode);
  * all S latent samples of a chunk go through ONE decoder GEMM whose output is written transposed and cell-major
    ([G, cells * S]: the S samples of a (cell, gene) are contiguous);
  * the noise-count log-probability is evaluated only at stored counts, with the sample average kept in registers
    (csrc/posterior.cu); cells are visited in barcode order, so the COO comes out globally (m, c)-sorted with no sort,
    and there is no host synchronisation between chunks (output rows are sized from the stored-count numbers);
  * the estimators consume the device-resident COO directly (estimation.DevicePosteriorCOO).
Monte-Carlo noise: torch's CUDA generator instead of pyro's trace order, so individual draws differ from a pyro run;
the distribution they are drawn from is the guide's (model.py:506-567), and y is fixed to its MAP (y_map=True).
"""
import logging
from typing import Dict, Optional

import numpy as np
import scipy.sparse as sp
import torch

from . import consts, ops
from .estimation import DevicePosteriorCOO, EstimationMethod, IndexConverter
from .sparse_utils import csr_set_rows_to_zero, subtract_noise_in_cells  # noqa: F401  (re-exported like the reference)
from .model import RemoveBackgroundPyroModel
from .vae import EncoderInputs

logger = logging.getLogger("cellbender")


class Posterior:
    """Posterior on latent variables and denoised counts (reference posterior.py:170-237)."""

    def __init__(self, dataset_obj, vi_model: Optional[RemoveBackgroundPyroModel], posterior_batch_size: int = 128,
                 counts_dtype: np.dtype = np.dtype(np.uint32), float_threshold: Optional[float] = 0.5, debug: bool = False,
                 cells_per_chunk: int = 512):
        self.dataset_obj = dataset_obj
        self.vi_model = vi_model
        if vi_model is not None:
            vi_model.eval()
            vi_model.encoder["z"].eval()
            vi_model.encoder["other"].eval()
            vi_model.decoder.eval()
        self.use_cuda = True
        self.device = "cuda" if vi_model is None else vi_model.device
        self.analyzed_gene_inds = None if dataset_obj is None else dataset_obj.analyzed_gene_inds
        self.count_matrix_shape = None if dataset_obj is None else dataset_obj.matrix.shape
        self.barcode_inds = None if self.count_matrix_shape is None else np.arange(0, self.count_matrix_shape[0])
        self.dtype = counts_dtype
        self.debug = debug
        self.float_threshold = float_threshold
        self.posterior_batch_size = posterior_batch_size
        self.cells_per_chunk = cells_per_chunk
        self.use_cuda_graph = True  # replay one captured graph per full chunk of the posterior pass (device_posterior)
        # pipelined pass: tile configuration of the decoder product (3 = 256 x 256 tiles, one CTA per SM) and residency cap
        # of the window kernel that runs beside it (tools/posterior_profile.py --sweep measures the alternatives)
        self.pipeline_tile_cfg = 3
        self.pipeline_window_ctas = 2
        self._noise_count_posterior_coo: Optional[sp.coo_matrix] = None
        self._noise_count_posterior_kwargs: Optional[dict] = None
        self._noise_count_posterior_coo_offsets: Optional[Dict[int, int]] = None
        self._noise_count_regularized_posterior_coo = None
        self._noise_count_regularized_posterior_kwargs = None
        self._device_posterior: Optional[DevicePosteriorCOO] = None
        self._latents: Optional[Dict[str, np.ndarray]] = None
        self._csr: Optional[ops.DeviceCSR] = None
        if dataset_obj is not None:
            self.index_converter = IndexConverter(total_n_cells=self.count_matrix_shape[0],
                                                  total_n_genes=self.count_matrix_shape[1])

    # ---- device-resident analysed count matrix -------------------------------------------------------------
    def _device_csr(self) -> ops.DeviceCSR:
        if self._csr is None:
            self._csr = ops.DeviceCSR.from_scipy(self.dataset_obj.get_count_matrix(), self.device)
        return self._csr

    def _amb_unit(self) -> torch.Tensor:
        chi = self.vi_model.param_store["chi_ambient"].detach()
        return (chi / torch.linalg.vector_norm(chi)).contiguous()

    def _encode(self, rows: torch.Tensor):
        csr = self._device_csr()
        out = ops.gather_rows(csr, rows, self._amb_unit(), want=("norm", "lognorm"))
        inp = EncoderInputs(out["norm"], out["lognorm"], out["feats"], csr.n_genes)
        m = self.vi_model
        enc = m.encoder(x=inp, chi_ambient=m.param_store["chi_ambient"].detach(), cell_prior_log=m.d_cell_loc_prior)
        return inp, enc

    # ---- latents (posterior.py:861-917) ----------------------------------------------------------------------
    @property
    def latents_map(self) -> Dict[str, np.ndarray]:
        if self._latents is None:
            self._get_latents_map()
        return self._latents

    @torch.no_grad()
    def _get_latents_map(self):
        if self.vi_model is None:
            self._latents = {"z": None, "d": None, "p": None, "phi_loc_scale": None, "epsilon": None}
            return None
        m = self.vi_model
        n = self._device_csr().shape[0]
        z = torch.empty((n, m.encoder["z"].output_dim), device=self.device)
        d, p, eps = (torch.empty(n, device=self.device) for _ in range(3))
        d_cell_scale = m.param_store["d_cell_scale"].reshape(())
        # The reference walks the droplets 500 at a time (posterior.py:868-871); the eval-mode encoders treat every droplet
        # on its own, so the batch size is free: 2048 droplets per (launch-bound) pass.  The means are the closed forms of
        # LogNormal(loc, scale).mean and Gamma(epsilon * c, c).mean (posterior.py:896-914).
        step = 2048
        for start in range(0, n, step):
            rows = torch.arange(start, min(n, start + step), device=self.device)
            _, enc = self._encode(rows)
            sl = slice(start, start + rows.numel())
            z[sl] = enc["z"]["loc"].reshape(rows.numel(), -1)
            d[sl] = torch.exp(enc["d_loc"] + 0.5 * d_cell_scale * d_cell_scale)
            p[sl] = enc["p_y"].sigmoid()
            eps[sl] = enc["epsilon"]
        self._latents = {"z": z.double().cpu().numpy(), "d": d.double().cpu().numpy(), "p": p.double().cpu().numpy(),
                         "phi_loc_scale": [m.param_store["phi_loc"].item(), m.param_store["phi_scale"].item()],
                         "epsilon": eps.double().cpu().numpy()}

    # ---- noise-count posterior (posterior.py:388-594, 669-859) -----------------------------------------------
    def cell_noise_count_posterior_coo(self, **kwargs) -> sp.coo_matrix:
        if (self._noise_count_posterior_coo is None) or (kwargs != self._noise_count_posterior_kwargs):
            self._get_cell_noise_count_posterior_coo(**kwargs)
            self._noise_count_posterior_kwargs = kwargs
        return self._noise_count_posterior_coo

    @torch.no_grad()
    def sample_rate_coefficients(self, enc: Dict[str, torch.Tensor], n_samples: int, y_map: bool = True,
                                 noise: Optional[Dict[str, torch.Tensor]] = None):
        """S latent draws from the guide for a chunk of B droplets (sample_mu_lambda_alpha, posterior.py:737-782, with the
        model's returned quantities in factored form: mu = a_mu * chi(z), lam = c_a * chi_ambient + c_b * chi_bar,
        alpha = 1 / phi), laid out CELL-MAJOR: z [B, S, Z]; a_mu, c_a, c_b [B, S]; alpha [S].
        NOTE the reference returns the lambda of the empty-droplet regulariser block (epsilon = 1, y = 0;
        model.py:385-393, 445), reproduced here.  ``noise`` (tests) supplies the draws themselves: phi [S]; rho, d_empty,
        d_cell, epsilon [B, S]; z [B, S, Z]."""
        if not y_map:
            raise NotImplementedError("only y_map=True (the value the reference's pipeline uses) is supported")
        m, ps = self.vi_model, self.vi_model.param_store
        B, S = enc["p_y"].shape[0], n_samples
        z_loc, z_scale = enc["z"]["loc"].reshape(B, -1), enc["z"]["scale"].reshape(B, -1)
        prob = enc["p_y"].sigmoid()
        if noise is None:
            phi_loc, phi_scale = ps["phi_loc"], ps["phi_scale"]
            # The guide's distributions (model.py:506-567) sampled from their definitions with torch's graph-capturable
            # generators (randn, _standard_gamma): Gamma(c, r) = G(c) / r, Beta(a, b) = G(a) / (G(a) + G(b)) (what
            # torch.distributions.Beta draws through a Dirichlet), LogNormal(l, s) = exp(l + s n).  torch.distributions
            # objects are avoided on purpose: their sampling path invalidates CUDA-graph capture (r02 probe).
            dev = self.device
            tiny = torch.finfo(torch.float32).tiny
            sg = lambda conc: torch._standard_gamma(conc.contiguous()).clamp_min(tiny)
            phi_loc, phi_scale = ps["phi_loc"].reshape(()), ps["phi_scale"].reshape(())
            phi = sg((phi_loc.pow(2) / phi_scale.pow(2)).expand(S)) / (phi_loc / phi_scale.pow(2))
            if m.include_rho:
                ga, gb = sg(ps["rho_alpha"].reshape(()).expand(B, S)), sg(ps["rho_beta"].reshape(()).expand(B, S))
                rho = ga / (ga + gb)
            else:
                rho = torch.zeros((B, S), device=dev)
            d_empty = torch.exp(ps["d_empty_loc"].reshape(()) + ps["d_empty_scale"].reshape(()) * torch.randn((B, S), device=dev))
            z = z_loc[:, None, :] + z_scale[:, None, :] * torch.randn((B, S, z_loc.shape[1]), device=dev)
            d_cell = torch.exp((prob * enc["d_loc"] + (1 - prob) * m.d_cell_loc_prior)[:, None]
                               + ps["d_cell_scale"].reshape(()) * torch.randn((B, S), device=dev))
            epsilon = sg(((prob * enc["epsilon"] + (1 - prob)) * m.epsilon_prior)[:, None].expand(B, S)) / m.epsilon_prior
        else:
            phi, rho, d_empty, z, d_cell, epsilon = (noise[k].to(self.device) for k in ("phi", "rho", "d_empty", "z", "d_cell", "epsilon"))
        y = (enc["p_y"] > 0).float()[:, None].expand(B, S)
        if m.model_type == "full":
            a_mu, c_a, c_b = (1 - rho) * y * epsilon * d_cell, (1 - rho) * d_empty, rho * d_empty
        elif m.model_type == "swapping":
            a_mu, c_a, c_b = (1 - rho) * y * epsilon * d_cell, torch.zeros_like(rho), rho * d_empty
        else:  # ambient
            a_mu, c_a, c_b = y * epsilon * d_cell, d_empty, torch.zeros_like(d_empty)
        return z, a_mu, c_a, c_b, phi.reciprocal()

    @torch.no_grad()
    def _get_cell_noise_count_posterior_coo(self, n_samples: int = 20, y_map: bool = True, n_counts_max: int = 20,
                                            smallest_log_probability: float = -10.0) -> sp.coo_matrix:
        dev_post = self.device_posterior(n_samples=n_samples, y_map=y_map, n_counts_max=n_counts_max,
                                         smallest_log_probability=smallest_log_probability)
        shape = [int(np.prod(self.count_matrix_shape, dtype=np.uint64)), n_counts_max]
        self._noise_count_posterior_coo = dev_post.to_scipy(shape)
        self._noise_count_posterior_coo_offsets = dev_post.offsets_dict()
        return self._noise_count_posterior_coo

    def posterior_cell_order(self, n_counts_max: int = 20):
        """The cells the reference's posterior loop visits (p > 0.5, minus a dropped trailing minibatch of < 4 cells,
        dataprep.py:155-161), re-ordered by barcode index of the full matrix so that the emitted COO is globally (m, c)
        sorted, with the per-cell window size of the reference's 128-cell minibatch the cell belongs to
        (n = min(n_counts_max, data.max()), posterior.py:816).  Returns (rows int64 [n] of the analysed matrix,
        n_window int32 [n] device tensor, barcodes int64 [n])."""
        csr = self._device_csr()
        cell_logic = self.latents_map["p"] > consts.CELL_PROB_CUTOFF
        if cell_logic.sum() == 0:
            logger.error(f"ERROR: Found zero droplets with posterior cell probability > {consts.CELL_PROB_CUTOFF}.")
            raise RuntimeError("Zero cells found!")
        cell_rows_all = np.flatnonzero(cell_logic)  # rows of the analysed count matrix, the reference's loader order
        feats = []
        for start in range(0, cell_rows_all.size, 4096):
            rows = torch.as_tensor(cell_rows_all[start:start + 4096], device=self.device)
            feats.append(ops.gather_feats(csr, rows)[:, 3])
        row_max = torch.cat(feats)
        pb = self.posterior_batch_size
        n_b = (cell_rows_all.size + pb - 1) // pb
        padded = torch.zeros(n_b * pb, device=self.device)
        padded[: row_max.numel()] = row_max
        group_max = padded.view(n_b, pb).max(dim=1).values
        tail = cell_rows_all.size % pb
        n_eff = cell_rows_all.size if (tail == 0 or tail >= consts.SMALLEST_ALLOWED_BATCH) else cell_rows_all.size - tail
        n_window_all = torch.clamp(group_max.repeat_interleave(pb)[:n_eff], max=float(n_counts_max)).to(torch.int32)
        barcodes = np.asarray(self.dataset_obj.analyzed_barcode_inds)[cell_rows_all[:n_eff]].astype(np.int64)
        order = np.argsort(barcodes, kind="stable")
        return cell_rows_all[:n_eff][order], n_window_all[torch.as_tensor(order, device=self.device)].contiguous(), barcodes[order]

    @torch.no_grad()
    def device_posterior(self, n_samples: int = 20, y_map: bool = True, n_counts_max: int = 20,
                         smallest_log_probability: float = -10.0, cell_subset: Optional[np.ndarray] = None,
                         lambda_multiplier: float = 1.0, seed: Optional[int] = None,
                         noise_fn=None) -> DevicePosteriorCOO:
        """The posterior COO, computed and left on the device, sorted by (m, c).

        ``cell_subset``: a contiguous range of positions in ``posterior_cell_order()`` (barcode-sorted cell list) --
        barcode sharding across GPUs: the shards' COO pieces concatenate, in rank order, to the sorted whole.
        ``seed``: reseed the generator per chunk with seed + position of the chunk's first cell, which makes the
        Monte-Carlo draws a function of the chunk only (sharded and unsharded runs then agree bit for bit when shard
        boundaries fall on chunk boundaries).  ``noise_fn(position_slice, enc) -> dict`` (tests) injects the latent draws.
        No host synchronisation happens between chunks; the single one is the read of the output sizes at the end."""
        m = self.vi_model
        csr = self._device_csr()
        G = csr.n_genes
        rows_sorted, n_window, barcodes = self.posterior_cell_order(n_counts_max)
        lo, hi = 0, rows_sorted.size
        if cell_subset is not None:
            cs = np.asarray(cell_subset)
            cs = cs[cs < rows_sorted.size]
            if cs.size:
                assert np.all(np.diff(cs) == 1), "cell_subset must be a contiguous range of the barcode-sorted cell list"
                lo, hi = int(cs[0]), int(cs[-1]) + 1
            else:
                lo = hi = 0
        S = n_samples
        # stored counts per cell from the host copy of the analysed matrix: output rows of every chunk are known up front
        host_indptr = self._host_indptr()
        nnz_cell = (host_indptr[rows_sorted[lo:hi] + 1] - host_indptr[rows_sorted[lo:hi]]).astype(np.int64)
        prefix = np.concatenate(([0], np.cumsum(nnz_cell)))
        ws = self._workspace(int(prefix[-1]), n_counts_max)
        rows_dev = torch.as_tensor(rows_sorted[lo:hi], device=self.device)
        barcode_dev = torch.as_tensor(barcodes[lo:hi], device=self.device)
        prefix_dev = torch.as_tensor(prefix, device=self.device)
        n_window = n_window[lo:hi]
        # per-model constants in persistent buffers (a cached chunk graph refers to them by address): refreshed in place
        pc = getattr(self, "_pc", None)
        if pc is None:
            ld = ops.row_pitch(G)
            pc = self._pc = {"chi_a": torch.zeros(ld, device=self.device), "chi_b": torch.zeros(ld, device=self.device),
                             "gene_all": torch.as_tensor(np.asarray(self.analyzed_gene_inds).astype(np.int64), device=self.device)}
        chi_a, chi_b, gene_all = pc["chi_a"], pc["chi_b"], pc["gene_all"]
        chi_a[:G].copy_(m.param_store["chi_ambient"].detach())
        chi_b[:G].copy_(m.avg_gene_expression)
        lin = m.decoder.out_linear
        cpc = self.cells_per_chunk
        n_sel = hi - lo
        total_genes = int(self.count_matrix_shape[1])
        H = int(lin.weight.shape[1])
        slots = self._pipeline_slots(G, S, cpc, H)
        for sl in slots:
            sl["estart"].zero_()  # a stage that runs ahead of / behind the chunk sequence sees an empty chunk

        # One chunk = three stages that hand over through the static buffers of a slot:
        #   stage 0  encoders + S latent draws + decoder hidden layer (a few hundred small launches, latency-bound)
        #   stage 1  ONE decoder GEMM for all samples of the chunk, output transposed and cell-major:
        #            logits_t[g, b * S + s] = W[g, :] . h[b * S + s, :]                                  (tensor-bound)
        #   stage 2  column logsumexp of logits_t + bias (HBM-bound), then the noise-count window over the chunk's stored
        #            counts (XU-bound)
        def stage0(sl, Bc, pos0):
            _, enc = self._encode(sl["rows"][:Bc])
            noise = None if noise_fn is None else noise_fn(slice(pos0, pos0 + Bc), enc)
            z, a_mu, c_a, c_b, alpha = self.sample_rate_coefficients(enc, S, y_map=y_map, noise=noise)
            h = m.decoder.hidden(z.reshape(Bc * S, -1))
            sl["h"][: Bc * S].copy_(h)
            sl["a_mu"][: Bc * S].copy_(a_mu.reshape(-1)), sl["c_a"][: Bc * S].copy_(c_a.reshape(-1))
            sl["c_b"][: Bc * S].copy_(c_b.reshape(-1)), sl["alpha"].copy_(alpha)

        def stage1(sl, Bc, tile_cfg=0):
            ops.gemm_tf32(lin.weight, sl["h"][: Bc * S], G, Bc * S, H, out=sl["logits"][:, : Bc * S], split_k=1, tile_cfg=tile_cfg)

        def stage2(sl, Bc, n_entries, ctas_per_sm=0):
            ops.col_logsumexp(sl["logits"], G, Bc * S, row_add=lin.bias, out=sl["lse"][: Bc * S])
            ops.posterior_window(
                csr, sl["rows"][:Bc], sl["logits"], lin.bias, sl["lse"], chi_a, chi_b, sl["a_mu"], sl["c_a"], sl["c_b"], sl["alpha"],
                sl["nwin"][:Bc], sl["estart"][: Bc + 1], n_entries, sl["barcode"][:Bc], gene_all, total_genes, ws, sl["offset"],
                log_thresh=smallest_log_probability, lambda_multiplier=lambda_multiplier, ctas_per_sm=ctas_per_sm)

        def load(sl, a0, a1):
            """Static operands of the chunk at positions [a0, a1) of the selected cell list."""
            n = a1 - a0
            sl["rows"][:n].copy_(rows_dev[a0:a1]), sl["nwin"][:n].copy_(n_window[a0:a1]), sl["barcode"][:n].copy_(barcode_dev[a0:a1])
            torch.sub(prefix_dev[a0:a1 + 1], prefix_dev[a0], out=sl["estart"][: n + 1])
            sl["offset"].copy_(prefix_dev[a0:a0 + 1])

        n_full = n_sel // cpc
        # Full chunks run as a three-deep software pipeline: iteration t replays ONE captured CUDA graph whose three
        # parallel branches are stage 0 of chunk t, stage 1 of chunk t - 1 and stage 2 of chunk t - 2 (slot = chunk mod 3,
        # so there are three graph variants): the latency-bound small launches hide under the two bulk kernels of the
        # neighbouring chunks.  The decoder product and the window kernel themselves contend for issue slots and gain
        # little from running side by side (r02ag timeline: 1.39 ms together, 0.66 + 0.70 ms alone); a fourth stage for the
        # logsumexp measured slower (36.2 vs 34.8 ms per pass).  Two extra iterations drain the pipeline; the stages that
        # have no chunk there see an empty one.  Seeded / injected-noise passes (tests, sharded-equals-unsharded checks) and
        # the ragged last chunk run the stages eagerly, one after the other.
        NS = self.N_PIPELINE_SLOTS
        use_graph = self.use_cuda_graph and seed is None and noise_fn is None and n_full >= 3
        if use_graph:
            chunk_entries = np.array([int(prefix[(i + 1) * cpc] - prefix[i * cpc]) for i in range(n_full)])
            key = (S, n_counts_max, float(smallest_log_probability), float(lambda_multiplier), bool(y_map), cpc)
            cached = getattr(self, "_chunk_graph", None)
            if not (cached is not None and cached["key"] == key and cached["ws"] is ws and cached["slots"] is slots
                    and cached["max_entries"] >= int(chunk_entries.max())):
                max_entries = int(chunk_entries.max() * 1.25) + 1024  # head-room: later passes reuse the captured graphs
                # the small launches of stage 0 outrank the two bulk kernels (priorities are recorded into the graph's kernel
                # nodes): at equal priority a small kernel waits until every pending CTA of a large grid launched before it
                # has been placed, and the stages run one after the other (r02ae timeline: no overlap at all)
                # The window kernel (XU-bound, no shared memory) is capped at two CTAs per SM and outranks the decoder product
                # (tensor-bound, one 256 x 256-tile CTA per SM): both stay resident side by side for their whole duration.
                streams = [torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)]
                capture_stream = torch.cuda.Stream(priority=-1)

                def iteration(v):
                    main = torch.cuda.current_stream()
                    for st_ in streams:
                        st_.wait_stream(main)
                    stage0(slots[v], cpc, 0)
                    with torch.cuda.stream(streams[0]):
                        stage1(slots[(v - 1) % NS], cpc, tile_cfg=self.pipeline_tile_cfg)
                    with torch.cuda.stream(streams[1]):
                        stage2(slots[(v - 2) % NS], cpc, max_entries, ctas_per_sm=self.pipeline_window_ctas)
                    for st_ in streams:
                        main.wait_stream(st_)

                graphs = []
                for v in range(NS):
                    capture_stream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(capture_stream):
                        iteration(v)  # warm-up outside capture (allocator, lazy initialisation) on empty chunks
                    torch.cuda.current_stream().wait_stream(capture_stream)
                    g_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_, stream=capture_stream):
                        iteration(v)
                    graphs.append(g_)
                cached = self._chunk_graph = {"key": key, "graphs": graphs, "max_entries": max_entries, "ws": ws, "slots": slots,
                                              "streams": streams}
            graphs = cached["graphs"]
            for t in range(n_full + NS - 1):
                if t < n_full:
                    load(slots[t % NS], t * cpc, (t + 1) * cpc)
                graphs[t % NS].replay()
            first_eager = n_full * cpc
        elif self.use_cuda_graph and seed is None and noise_fn is None and n_full >= 1:
            # one or two full chunks (small barcode shards): the pipeline's two drain iterations would cost more than they
            # hide -- the three stages of a chunk as ONE captured graph, replayed per chunk
            chunk_entries = np.array([int(prefix[(i + 1) * cpc] - prefix[i * cpc]) for i in range(n_full)])
            key = (S, n_counts_max, float(smallest_log_probability), float(lambda_multiplier), bool(y_map), cpc)
            cached = getattr(self, "_chunk_graph_serial", None)
            if not (cached is not None and cached["key"] == key and cached["ws"] is ws and cached["slots"] is slots
                    and cached["max_entries"] >= int(chunk_entries.max())):
                max_entries = int(chunk_entries.max() * 1.25) + 1024

                def whole_chunk():
                    stage0(slots[0], cpc, 0)
                    stage1(slots[0], cpc)
                    stage2(slots[0], cpc, max_entries)

                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    whole_chunk()  # warm-up outside capture on an empty chunk
                torch.cuda.current_stream().wait_stream(side)
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    whole_chunk()
                cached = self._chunk_graph_serial = {"key": key, "graph": g_, "max_entries": max_entries, "ws": ws, "slots": slots}
            for t in range(n_full):
                load(slots[0], t * cpc, (t + 1) * cpc)
                cached["graph"].replay()
            first_eager = n_full * cpc
        else:
            first_eager = 0
        for start in range(first_eager, n_sel, cpc):
            stop = min(n_sel, start + cpc)
            if seed is not None:
                torch.manual_seed(int(seed) + lo + start)
            sl = slots[0]
            load(sl, start, stop)
            stage0(sl, stop - start, lo + start)
            stage1(sl, stop - start)
            stage2(sl, stop - start, int(prefix[stop] - prefix[start]))
        post = DevicePosteriorCOO(*ops.posterior_compact(ws, smallest_log_probability), n_cols=n_counts_max)
        if cell_subset is None:
            self._device_posterior = post
        return post

    def _workspace(self, n_entries: int, n_counts_max: int) -> "ops.PosteriorWorkspace":
        """Per-stored-count scratch of a pass, kept between passes (grown when needed) so that the cached chunk graph,
        which refers to it by address, stays valid; the compaction at the end of a pass copies the result out."""
        ws = getattr(self, "_ws", None)
        if ws is None or ws.n_counts_max != n_counts_max or ws.capacity < n_entries:
            ws = self._ws = ops.PosteriorWorkspace(int(n_entries * 1.1) + 1024, n_counts_max, self.device)
        ws.n_entries = int(n_entries)
        return ws

    N_PIPELINE_SLOTS = 3  # stages of a chunk = depth of the software pipeline of ``device_posterior``

    def _pipeline_slots(self, G: int, S: int, cpc: int, H: int):
        """``N_PIPELINE_SLOTS`` sets of static buffers through which the stages of a chunk hand over (see ``device_posterior``); kept
        between passes: the captured graphs refer to them by address.  logits: [G, cells * S] transposed, cell-major
        (1.37 GB per slot at the PBMC8k shape)."""
        key = (G, S, cpc, H)
        if getattr(self, "_slots_key", None) != key:
            dev = self.device
            f = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
            self._slots = [{"logits": torch.empty((G, ops.row_pitch(S * cpc)), device=dev), "lse": f(S * cpc), "h": f(cpc * S, H),
                            "a_mu": f(cpc * S), "c_a": f(cpc * S), "c_b": f(cpc * S), "alpha": f(S),
                            "rows": torch.zeros(cpc, dtype=torch.int64, device=dev),
                            "nwin": torch.zeros(cpc, dtype=torch.int32, device=dev),
                            "estart": torch.zeros(cpc + 1, dtype=torch.int64, device=dev),
                            "barcode": torch.zeros(cpc, dtype=torch.int64, device=dev),
                            "offset": torch.zeros(1, dtype=torch.int64, device=dev)} for _ in range(self.N_PIPELINE_SLOTS)]
            self._slots_key = key
            self._chunk_graph = self._chunk_graph_serial = None
        return self._slots

    def _host_indptr(self) -> np.ndarray:
        if getattr(self, "_indptr_host", None) is None:
            self._indptr_host = self._device_csr().indptr.cpu().numpy()
        return self._indptr_host

    # ---- posterior regularisation (posterior.py:332-378) ----------------------------------------------------
    def regularize_posterior(self, regularization, **kwargs) -> sp.coo_matrix:
        """Do posterior regularization (``regularization`` = regularization.PRq); caches like the reference."""
        cached = self._noise_count_regularized_posterior_kwargs
        if cached is not None:
            same = all((v == regularization.name()) if k == "method" else (k in kwargs and kwargs[k] == v)
                       for k, v in cached.items())
            if same:
                logger.debug("Regularized posterior is already cached")
                return self._noise_count_regularized_posterior_coo
        if self._noise_count_posterior_coo is None:
            self.cell_noise_count_posterior_coo()
        self._noise_count_regularized_posterior_coo = regularization.regularize(
            noise_count_posterior_coo=self._noise_count_posterior_coo,
            noise_offsets=self._noise_count_posterior_coo_offsets, index_converter=self.index_converter, **kwargs)
        kwargs.update({"method": regularization.name()})
        kwargs.pop("raw_count_matrix", None)
        self._noise_count_regularized_posterior_kwargs = kwargs
        return self._noise_count_regularized_posterior_coo

    # ---- denoised counts (posterior.py:282-330) --------------------------------------------------------------
    def compute_denoised_counts(self, estimator_constructor: "type[EstimationMethod]", **kwargs) -> sp.csc_matrix:
        """Clean output count matrix: observed counts of the called cells minus the point estimate of their noise.
        A regularized posterior, once computed, takes priority (posterior.py:297-303)."""
        estimator = estimator_constructor(index_converter=self.index_converter)
        if self._noise_count_regularized_posterior_coo is not None:
            logger.debug("Using regularized posterior to compute denoised counts")
            noise_csr = estimator.estimate_noise(noise_log_prob_coo=self._noise_count_regularized_posterior_coo,
                                                 noise_offsets=self._noise_count_posterior_coo_offsets, **kwargs)
        else:
            if self._noise_count_posterior_coo is None and self._device_posterior is None:
                self.cell_noise_count_posterior_coo()
            if self._device_posterior is not None:
                noise_csr = estimator.estimate_noise(noise_log_prob_coo=self._device_posterior, noise_offsets=None, **kwargs)
            else:
                noise_csr = estimator.estimate_noise(noise_log_prob_coo=self._noise_count_posterior_coo,
                                                     noise_offsets=self._noise_count_posterior_coo_offsets, **kwargs)
        # subtract cell noise from the observed counts of the called cells; every other droplet is empty in the output
        # (posterior.py:315-328), one fused device pass (csrc/csr_combine.cu)
        count_matrix = self.dataset_obj.matrix  # all barcodes
        cell_inds = np.asarray(self.dataset_obj.analyzed_barcode_inds)[self.latents_map["p"] > consts.CELL_PROB_CUTOFF]
        return subtract_noise_in_cells(count_matrix, noise_csr, cell_inds)
