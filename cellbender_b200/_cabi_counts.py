"""Kernel launches per C-ABI call (for bench.py's ``gpu_launches`` claim).  Conservative: a split-K product counts as
one launch although its fixed-order reduction is a second kernel; a range sweep counts one although ``advance`` adds the
one-thread counter kernel.  tests/test_cabi.py checks the table against include/cellbender_b200.h."""
LAUNCHES = {
    # training step
    "cb_csr_gather_rows": 1, "cb_encoder_features": 1, "cb_bn0_forward": 1, "cb_bn0_backward": 1, "cb_bn_act_forward": 1,
    "cb_bn_act_backward": 1, "cb_head_forward": 1, "cb_head_backward": 1, "cb_gemm_tf32": 1, "cb_gemm_tf32_pair": 1,
    "cb_gemm_tf32_ex": 1, "cb_feat_dw": 1, "cb_latent_prepare": 1, "cb_latent_beta_sample": 1, "cb_latent_forward": 1,
    "cb_latent_backward": 1, "cb_simplex_forward": 1, "cb_simplex_backward": 1, "cb_elbo_obs_fwd_bwd": 3, "cb_elbo_obs_exact_fwd_bwd": 3, "cb_colsum_rows": 1,
    "cb_obs_loss": 1, "cb_obs_small_grads": 1, "cb_clipped_adam_step": 2, "cb_clipped_adam_range": 1, "cb_clipped_adam_p2p": 2,
    "cb_clipped_adam_nvls": 2, "cb_clipped_adam_nvls_ex": 2,
    # elementwise likelihoods (the reference's distribution classes)
    "cb_nbpc_approx_log_prob_fwd": 1, "cb_nbpc_approx_log_prob_bwd": 1, "cb_nbpc_exact_log_prob_fwd": 1,
    "cb_nbpc_exact_log_prob_bwd": 1,
    # posterior
    "cb_col_logsumexp": 2, "cb_posterior_noise_window": 1, "cb_posterior_offset_flags": 1, "cb_posterior_compact": 1,
    "cb_dense_nonzero_flags": 1, "cb_dense_to_coo": 1, "cb_prq_regularize": 1, "cb_posterior_tilt": 1,
    # estimation
    "cb_exclusive_scan_i32": 3, "cb_segment_heads": 1, "cb_segment_starts": 1, "cb_segment_estimate": 1,
    "cb_csr_from_segments": 1, "cb_mckp_gene_deficit": 2, "cb_mckp_candidates": 1, "cb_mckp_compact": 1, "cb_mckp_select": 1,
    "cb_mckp_finalize": 1,
    # output assembly
    "cb_csr_combine_flags": 1, "cb_csr_combine_fill": 1, "cb_csr_to_csc": 2,
}
# entry points that only answer a question on the host (no launch)
QUERIES = ("cb_abi_version", "cb_gemm_tf32_workspace_elems", "cb_gemm_tf32_suggest_split", "cb_latent_num_consts",
           "cb_latent_num_terms", "cb_posterior_window_stride", "cb_col_logsumexp_scratch_elems", "cb_scan_scratch_elems",
           "cb_elbo_obs_scratch_elems")
