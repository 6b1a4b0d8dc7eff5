"""-m gpu parity tests of the output assembly (csrc/csr_combine.cu through cellbender_b200.sparse_utils): bit-exact
against the unmodified reference's outputs (tests/golden/sparse_utils.npz), against the oracle on random matrices of
every value type, and through size-independent properties at PBMC8k size.  Mirrors the reference's
tests/test_sparse_utils.py:128-196."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import sparse_utils as osp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def su():
    from cellbender_b200 import build, sparse_utils

    build.build()
    return sparse_utils


def _canonical(m) -> bool:
    m = m.copy()
    ok = m.has_sorted_indices or True
    ind, ptr = m.indices, m.indptr
    for k in range(len(ptr) - 1):
        seg = ind[ptr[k]:ptr[k + 1]]
        ok = ok and bool(np.all(np.diff(seg) > 0))
    return ok and bool(np.all(m.data != 0))


def _same(got, want_dense, fmt, nnz=None):
    assert got.format == fmt
    np.testing.assert_array_equal(np.asarray(got.todense()), want_dense)
    assert _canonical(got), "indices must be sorted, unique, without explicit zeros"
    assert got.nnz == (int(np.count_nonzero(want_dense)) if nnz is None else int(nnz))


def test_overwrite_and_row_zeroing_match_reference_golden(su, golden_dir):
    g = np.load(os.path.join(golden_dir, "sparse_utils.npz"))
    for tag in ("t0", "t1", "t2", "r0", "r1", "none"):
        m1, m2 = sp.csc_matrix(g[f"ow_{tag}_m1"]), sp.csc_matrix(g[f"ow_{tag}_m2"])
        keep1, keep2 = m1.copy(), m2.copy()
        out = su.overwrite_matrix_with_columns_from_another(mat1=m1, mat2=m2, column_inds=g[f"ow_{tag}_cols"])
        _same(out, g[f"ow_{tag}_out"], "csc")
        assert (m1 != keep1).nnz == 0 and (m2 != keep2).nnz == 0, "inputs must not be modified"
        # the reference test's own assertions (tests/test_sparse_utils.py:141-167)
        cols = list(g[f"ow_{tag}_cols"])
        excluded = [i for i in range(m1.shape[1]) if i not in cols]
        assert (out[:, excluded] != m2[:, excluded]).nnz == 0
        assert (out[:, cols] != m1[:, cols]).nnz == 0
    for tag in ("t0", "t1", "r"):
        m = sp.csr_matrix(g[f"rz_{tag}_m"])
        out = su.csr_set_rows_to_zero(m, list(g[f"rz_{tag}_rows"]))
        _same(out, g[f"rz_{tag}_out"], "csr", g[f"rz_{tag}_nnz"])
        assert out is m, "in place for a csr_matrix input, like the reference"
        out2 = su.csr_set_rows_to_zero(sp.csc_matrix(g[f"rz_{tag}_m"]), set(int(r) for r in g[f"rz_{tag}_rows"]))
        _same(out2, g[f"rz_{tag}_out"], "csr")
    with pytest.raises(ValueError):
        su.csr_set_rows_to_zero(np.zeros((2, 2)), [0])


def test_denoise_and_restore_match_reference_golden(su, golden_dir):
    g = np.load(os.path.join(golden_dir, "sparse_utils.npz"))
    cell_inds = g["dn_bcs"][g["dn_p"] > 0.5]
    count = sp.csr_matrix(g["dn_raw"])
    before = count.copy()
    den = su.subtract_noise_in_cells(count, sp.csr_matrix(g["dn_noise"]), cell_inds)
    _same(den, g["dn_denoised"], "csc", g["dn_denoised_nnz"])
    assert (count != before).nnz == 0
    res = su.restore_eliminated_features_in_cells(den, sp.csc_matrix(g["dn_raw"]), g["dn_genes"], g["dn_bcs"], g["dn_p"])
    _same(res, g["dn_restored"], "csc", g["dn_restored_nnz"])
    # PreparedDataset method = SingleCellRNACountsDataset.restore_eliminated_features_in_cells
    from cellbender_b200.synth import PreparedDataset

    ds = PreparedDataset(matrix=sp.csr_matrix(g["dn_raw"]), analyzed_barcode_inds=g["dn_bcs"], empty_barcode_inds=np.array([]),
                         analyzed_gene_inds=g["dn_genes"], priors={}, empty_UMI_threshold=0.0)
    _same(ds.restore_eliminated_features_in_cells(den, g["dn_p"]), g["dn_restored"], "csc")


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.uint32, np.float32, np.float64, np.int16])
@pytest.mark.parametrize("sign", [1, -1])
def test_combine_random_vs_oracle_all_value_types(su, dtype, sign):
    rng = np.random.default_rng(np.dtype(dtype).itemsize * 10 + sign)
    R, G = 257, 1031  # rows of very different lengths; B-only, A-only and shared columns; cancellations
    a = (rng.poisson(2.0, size=(R, G)) * (rng.random((R, G)) < 0.08)).astype(dtype)
    a[5] = 0  # empty row
    a[7] = rng.poisson(3.0, size=G)  # full row (longer than a warp's stride many times)
    b = (rng.poisson(2.0, size=(R, G)) * (rng.random((R, G)) < 0.08)).astype(dtype)
    b[rng.random((R, G)) < 0.02] = 0
    shared = rng.random((R, G)) < 0.03
    b[shared] = a[shared]  # a - b cancels to an explicit zero that must be dropped
    keep_rows = np.flatnonzero(rng.random(R) < 0.7)
    cols = np.flatnonzero(rng.random(G) < 0.6)
    A, B = sp.csr_matrix(a), sp.csr_matrix(b)
    # plain A + sign * B with a row mask
    want = np.where(np.isin(np.arange(R), keep_rows)[:, None], a.astype(np.float64) + sign * b.astype(np.float64), 0)
    got = su.combine_csr(A, B, sign=sign, keep_rows=keep_rows)
    _same(got, want.astype(got.dtype), "csr")
    assert got.dtype == su._value_dtype(np.dtype(dtype))
    if sign == 1:
        _same(su.overwrite_matrix_with_columns_from_another(A, B, cols),
              np.asarray(osp.overwrite_matrix_with_columns_from_another(A, B, cols).todense()), "csc")
    else:  # (the oracle's dense arithmetic in the kernels' value type: unsigned inputs would wrap around in numpy)
        wide = su._value_dtype(np.dtype(dtype))
        _same(su.subtract_noise_in_cells(A, B, keep_rows),
              np.asarray(osp.denoised_counts(A.astype(wide), B.astype(wide), keep_rows).todense()), "csc")


def test_combine_edge_cases(su):
    z = sp.csr_matrix((4, 6), dtype=np.int32)
    assert su.combine_csr(z, z).nnz == 0
    assert su.combine_csr(sp.csr_matrix((0, 5), dtype=np.int32)).shape == (0, 5)
    a = sp.csr_matrix(np.arange(24).reshape(4, 6).astype(np.int32))
    np.testing.assert_array_equal(su.combine_csr(a, None).toarray(), a.toarray())
    np.testing.assert_array_equal(su.combine_csr(a, z, fmt="csc").toarray(), a.toarray())
    np.testing.assert_array_equal(su.combine_csr(a, a, sign=-1).toarray(), 0 * a.toarray())
    np.testing.assert_array_equal(su.combine_csr(a, keep_rows=[]).toarray(), 0 * a.toarray())
    # non-canonical input (duplicates, unsorted indices) is canonicalised like scipy's binop does
    dup = sp.csr_matrix((np.array([1, 2, 3], dtype=np.int32), np.array([3, 1, 3]), np.array([0, 3, 3, 3, 3])), shape=(4, 6))
    np.testing.assert_array_equal(su.combine_csr(dup, a, keep_rows=[0]).toarray()[0], a.toarray()[0] + np.array([0, 2, 0, 4, 0, 0]))
    with pytest.raises(ValueError):
        su.combine_csr(a, sp.csr_matrix((4, 7), dtype=np.int32))
    with pytest.raises(TypeError):
        su.combine_csr(sp.csr_matrix(np.ones((2, 2), dtype=np.complex64)))


def test_output_assembly_full_size_properties(su):
    """PBMC8k-size: 75 000 barcodes x 33 538 features.  Size-independent properties: column sums of the selected /
    complementary columns, emptied rows, idempotence, and csr -> csc -> csr round trip (a checksum of checksums)."""
    gen = torch.Generator(device="cuda").manual_seed(0)
    R, G, per_row = 75_000, 33_538, 160
    rows = torch.arange(R, device="cuda").repeat_interleave(per_row)
    cols = torch.randint(0, G, (R * per_row,), device="cuda", generator=gen)
    key = torch.unique(rows * G + cols)  # sorted, duplicate-free
    r, c = (key // G).cpu().numpy(), (key % G).cpu().numpy()
    rng = np.random.default_rng(1)
    raw = sp.csr_matrix((rng.integers(1, 40, size=r.size).astype(np.int32), (r, c)), shape=(R, G))
    cell_inds = np.sort(rng.choice(R, 8000, replace=False))
    in_cell = np.isin(np.repeat(np.arange(R), np.diff(raw.indptr)), cell_inds)
    noise = raw.copy()  # noise counts exist only for the called cells (the posterior is computed for them alone)
    noise.data = (rng.binomial(raw.data, 0.3) * in_cell).astype(np.int32)
    noise.eliminate_zeros()
    den = su.subtract_noise_in_cells(raw, noise, cell_inds)
    assert den.format == "csc" and den.shape == (R, G) and _canonical(den)
    want_colsum = np.asarray(raw[cell_inds].sum(axis=0)).ravel() - np.asarray(noise[cell_inds].sum(axis=0)).ravel()
    np.testing.assert_array_equal(np.asarray(den.sum(axis=0)).ravel(), want_colsum)
    rowsum = np.asarray(den.sum(axis=1)).ravel()
    assert np.all(rowsum[np.setdiff1d(np.arange(R), cell_inds)] == 0)
    # idempotence of the row zeroing, and round trip through the device csc conversion
    again = su.combine_csr(den, keep_rows=cell_inds, fmt="csc")
    assert (again != den).nnz == 0
    back = su.combine_csr(den.tocsr(), fmt="csr")
    assert (back != den.tocsr()).nnz == 0
    genes = np.sort(rng.choice(G, 30_000, replace=False))
    p = np.zeros(R)
    p[cell_inds] = 0.9
    res = su.restore_eliminated_features_in_cells(den, raw, genes, np.arange(R), p)
    sel = np.zeros(G, dtype=bool)
    sel[genes] = True
    got_colsum = np.asarray(res.sum(axis=0)).ravel()
    raw_cells_colsum = np.asarray(raw[cell_inds].sum(axis=0)).ravel()
    np.testing.assert_array_equal(got_colsum[sel], want_colsum[sel])
    np.testing.assert_array_equal(got_colsum[~sel], raw_cells_colsum[~sel])


def test_dense_to_sparse_op_torch_matches_torch_nonzero(su):
    """reference sparse_utils.py:10-33 = torch.nonzero(as_tuple=True) + gather; all value types, padded rows,
    `tensor_for_nonzeros`, 1-d and 3-d inputs, empty results."""
    g = torch.Generator(device="cuda").manual_seed(3)
    for dtype in (torch.float32, torch.float64, torch.int32, torch.int64):
        t = (torch.randint(0, 4, (37, 101), device="cuda", generator=g) * (torch.rand(37, 101, device="cuda", generator=g) < 0.3))
        t = t.to(dtype)
        got = su.dense_to_sparse_op_torch(t)
        want = torch.nonzero(t, as_tuple=True)
        assert len(got) == 3 and got[0].dtype == torch.int64 and got[2].dtype == dtype
        for a, b in zip(got[:2], want):
            assert torch.equal(a, b)
        assert torch.equal(got[2], t[want].flatten())
    padded = torch.zeros(9, 104, device="cuda")
    padded[:, :101] = torch.rand(9, 101, device="cuda", generator=g).round()
    view = padded[:, :101]  # row stride 104: the layout of this package's dense minibatches
    got = su.dense_to_sparse_op_torch(view)
    want = torch.nonzero(view, as_tuple=True)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(got[2], view[want])
    other = torch.rand(9, 101, device="cuda", generator=g)
    got = su.dense_to_sparse_op_torch(other, tensor_for_nonzeros=view > 0)
    assert torch.equal(got[0], want[0]) and torch.equal(got[2], other[want])
    v = torch.tensor([0.0, 2.0, 0.0, 5.0], device="cuda")
    got = su.dense_to_sparse_op_torch(v)
    assert torch.equal(got[0], torch.tensor([1, 3], device="cuda")) and torch.equal(got[1], torch.tensor([2.0, 5.0], device="cuda"))
    cube = (torch.rand(3, 5, 7, device="cuda", generator=g) < 0.2).float()
    got = su.dense_to_sparse_op_torch(cube)
    want = torch.nonzero(cube, as_tuple=True)
    assert len(got) == 4 and all(torch.equal(a, b) for a, b in zip(got[:3], want))
    got = su.dense_to_sparse_op_torch(torch.zeros(4, 6, device="cuda"))
    assert all(x.numel() == 0 for x in got)
    with pytest.raises(RuntimeError):
        su.dense_to_sparse_op_torch(torch.zeros(4, 6))


def test_dataloader_roundtrip_through_dense_to_sparse_like_the_reference_tests(su):
    """The reference's tests/test_sparse_utils.py:47-93 and tests/test_dataprep.py:20-100: iterate the (unshuffled, and
    the sorted) DataLoader, sparsify every minibatch, un-sort the barcode indices, rebuild the count matrix."""
    from cellbender_b200.dataprep import DataLoader

    rng = np.random.default_rng(5)  # 55 rows = 11 full minibatches (a remainder below SMALLEST_ALLOWED_BATCH is dropped)
    matrix = sp.csr_matrix((rng.poisson(1.5, size=(55, 40)) * (rng.random((55, 40)) < 0.3)).astype(np.int32))
    loaders = [DataLoader(matrix, empty_drop_dataset=None, batch_size=5, fraction_empties=0.0, shuffle=False),
               DataLoader(matrix, empty_drop_dataset=None, batch_size=5, fraction_empties=0.0, shuffle=False,
                          sort_by=lambda x: -1 * np.array(x.max(axis=1).todense()).squeeze())]
    with pytest.raises(AssertionError):
        DataLoader(matrix, empty_drop_dataset=None, batch_size=5, fraction_empties=0.0, shuffle=True,
                   sort_by=lambda x: -1 * np.array(x.max(axis=1).todense()).squeeze())
    outs = []
    for loader in loaders:
        barcodes, genes, counts, ind = [], [], [], 0
        for data in loader:
            bcs_i_chunk, genes_i, counts_i = su.dense_to_sparse_op_torch(data)
            bcs_i = loader.unsort_inds(bcs_i_chunk + ind)
            barcodes.append(bcs_i.detach().cpu()), genes.append(genes_i.detach().cpu()), counts.append(counts_i.detach().cpu())
            ind += data.shape[0]
        out = sp.csc_matrix((np.concatenate(counts).astype(np.uint32),
                             (np.concatenate(barcodes).astype(np.uint32), np.concatenate(genes).astype(np.uint32))),
                            shape=matrix.shape)
        assert (out != matrix.tocsc()).nnz == 0
        outs.append(out)
    assert (outs[0] != outs[1]).nnz == 0
