"""CPU: the product's host logic end to end against the oracle, with the CUDA library replaced by the numpy test
double of tests/_host_double.py.  The test bodies are the GPU parity tests themselves (tests/test_gpu_parity.py,
tests/test_gpu_edges.py: same inputs, same tolerances); on the GPU box they run on the real kernels, here they pin
the planner, the launch tables, the GEMM group lists, the decomposition packing / rank rule / unpacking and the
coarse-graining drivers without a GPU."""
import numpy as np
import pytest

import _host_double
import test_gpu_edges as GE
import test_gpu_parity as GP
import test_z2_golden as Z2


@pytest.fixture
def gtn_host(monkeypatch):
    gtn, saved = _host_double.install(monkeypatch)
    yield gtn
    _host_double.uninstall(saved)


@pytest.mark.parametrize("trim", [False, True])
@pytest.mark.parametrize("case", range(len(GP.EINSUM_CASES)))
def test_einsum_vs_oracle(gtn_host, case, trim):
    GP.test_einsum_vs_oracle(gtn_host, case, trim)


def test_einsum_real_dtype(gtn_host):
    GP.test_einsum_real_dtype(gtn_host)


def test_format_encoder_switches_bit_exact(gtn_host):
    GP.test_format_encoder_switches_bit_exact(gtn_host)


@pytest.mark.parametrize("cut", [None, 8, 6])
def test_svd_vs_oracle(gtn_host, cut):
    GP.test_svd_vs_oracle(gtn_host, cut)


def test_svd_with_bosonic_legs(gtn_host):
    GP.test_svd_with_bosonic_legs(gtn_host)


def test_hconjugate_bit_exact(gtn_host):
    GP.test_hconjugate_bit_exact(gtn_host)


def test_eig_vs_oracle(gtn_host):
    GP.test_eig_vs_oracle(gtn_host)


@pytest.mark.parametrize("fmt", ["dense", "block"])
@pytest.mark.parametrize("algo", ["trg", "atrg2dy", "atrg2dx"])
def test_cg_step_vs_oracle(gtn_host, algo, fmt):
    GP.test_cg_step_vs_oracle(gtn_host, algo, fmt)


@pytest.mark.parametrize("case", range(len(GP.JOIN_CASES)))
def test_join_split_legs_bit_exact(gtn_host, case):
    GP.test_join_split_legs_bit_exact(gtn_host, case)


# ---- edge cases (tests/test_gpu_edges.py)
@pytest.mark.parametrize("case", range(len(GE.SMALL)))
def test_small_and_ragged_einsum(gtn_host, case):
    GE.test_small_and_ragged_einsum(gtn_host, case)


def test_einsum_errors(gtn_host):
    GE.test_einsum_errors(gtn_host)


def test_eig_rejects_non_hermitian(gtn_host):
    GE.test_eig_rejects_non_hermitian(gtn_host)


def test_svd_cutoff_larger_than_rank_and_full(gtn_host):
    GE.test_svd_cutoff_larger_than_rank_and_full(gtn_host)


def test_svd_of_zero_sector_and_rank_one(gtn_host):
    GE.test_svd_of_zero_sector_and_rank_one(gtn_host)


def test_block_non_power_of_two(gtn_host):
    GE.test_block_non_power_of_two(gtn_host)


def test_eig_reconstruction_D16(gtn_host):
    GE.test_eig_reconstruction_D16(gtn_host)


# ---- the product against the REAL reference's numbers on the Z2 gauge tensor (tests/test_z2_golden.py)
@pytest.mark.parametrize("name,fmt,algo,cut,steps", [c for c in Z2.GPU_CONFIGS if c[3] <= 16])
def test_host_vs_reference_on_z2(gtn_host, name, fmt, algo, cut, steps):
    Z2.test_gpu_vs_reference_on_z2(gtn_host, name, fmt, algo, cut, steps)


def test_host_hotrg3dz_random_vs_reference(gtn_host):
    Z2.test_gpu_hotrg3dz_random_vs_reference(gtn_host)


# ---- examples/example.py (the reference's example script on the product's API), log / checkpoint / resume
def test_example_script_against_reference_numbers(gtn_host, tmp_path, capsys):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "examples"))
    try:
        import example
    finally:
        sys.path.pop(0)
    ref = np.load(os.path.join(Z2.G, "z2_cg.npz"))["dense_atrg_chi8"]       # real reference: Tnorm, err, F per step
    log, ck = str(tmp_path / "run.jsonl"), str(tmp_path / "ck")
    recs = example.main(["--cgsteps", "3", "--Dcutxy", "8", "--log", log, "--checkpoint", ck])
    out = capsys.readouterr().out
    assert " _atrg:" in out and "_ini:" in out and "parameters:" in out
    assert len(recs) == 4
    for i, r in enumerate(recs):
        assert abs(r["F"] - complex(ref[i, 2], ref[i, 3])) <= 1e-10 * abs(r["F"]), (i, r["F"])
        if i:
            assert abs(r["Tnorm"] - ref[i, 0]) <= 1e-10 * ref[i, 0]
    # interrupted after 2 steps, resumed for the 3rd: same numbers
    ck2, log2 = str(tmp_path / "ck2"), str(tmp_path / "run2.jsonl")
    example.main(["--cgsteps", "2", "--Dcutxy", "8", "--log", log2, "--checkpoint", ck2])
    recs2 = example.main(["--cgsteps", "3", "--Dcutxy", "8", "--log", log2, "--checkpoint", ck2, "--resume"])
    assert len(recs2) == 4
    for a, b in zip(recs, recs2):
        assert abs(a["F"] - b["F"]) <= 1e-12 * abs(a["F"])
    from grassmanntn_b200 import checkpoint
    assert len(checkpoint.RunLog(log2).read()) == 4 and checkpoint.latest_step(ck2)[0] == 3
    # block format (example_block.py) and TRG
    recs3 = example.main(["--cgsteps", "2", "--Dcutxy", "16", "--trg", "--block"])
    refb = np.load(os.path.join(Z2.G, "z2_cg.npz"))["block_trg_chi32"]
    assert len(recs3) == 3 and abs(recs3[0]["F"] - complex(refb[0, 2], refb[0, 3])) <= 1e-10 * abs(recs3[0]["F"])
    with pytest.raises(SystemExit):
        example.main(["--beta", "2.0"])                                     # no fixture for other parameters


# ---- the truncated sector SVD (subspace iteration + certificate) on the host double
@pytest.fixture
def gtn_host_trunc(monkeypatch):
    gtn, saved = _host_double.install(monkeypatch, truncated=True)
    yield gtn
    _host_double.uninstall(saved)


@pytest.mark.parametrize("kind", ["decaying", "flat"])
def test_truncated_svd_path(gtn_host_trunc, kind):
    GP.test_truncated_svd_path(gtn_host_trunc, kind)


def test_vectorised_emulation_equals_scalar():
    """the vectorised kernel emulation of the test double against the scalar one (tests/test_tables_cpu.py) on a job
    with the full sign program (alpha, beta, Q cross terms between super-axes, conj)"""
    from test_tables_cpu import emulate
    from grassmanntn_b200 import _engine as E
    rng = np.random.RandomState(3)
    shape = (4, 2, 8, 4, 2)
    n = int(np.prod(shape))
    src = rng.rand(n) + 1j * rng.rand(n)
    in_st = E._row_strides(shape)
    perm = (3, 0, 4, 2, 1)
    out_shape = tuple(shape[a] for a in perm)
    ost = E._row_strides(out_shape)
    out_st = [ost[perm.index(a)] for a in range(5)]
    legs = []
    for a, d in enumerate(shape):
        pc = E._popcount_vec(np.arange(d))
        legs.append(E.LegTab(d, np.arange(d) * in_st[a], np.arange(d) * out_st[a], p=pc & 1, q=(pc >> 1) & 1))
    Q = [0] * 5
    for x, y in ((0, 3), (1, 2), (2, 4), (0, 4)):
        Q[x] |= 1 << y
        Q[y] |= 1 << x
    f, tabs = E.build_job(legs, alpha=[1, 0, 1, 0, 0], beta=[0, 1, 0, 0, 1], Q=Q, const=1, conj=True)
    a, b = np.zeros(n, complex), np.zeros(n, complex)
    emulate(f, tabs, src, a, 0.5)
    _host_double.emulate_vec(f, tabs, src, b, 0.5)
    assert len(tabs) >= 2 and np.array_equal(a, b)


def test_truncated_path_on_the_z2_chain_vs_reference(gtn_host_trunc):
    """three block-format TRG steps at chi = 32 on the Z2 gauge tensor: the third runs the truncated sector SVD
    (512 x 512 sector matrices, 48-row subspace) -- Tnorm, F and the trace error against the real reference"""
    from grassmanntn_b200 import _ops
    before = dict(_ops.SVD_PATH_STATS)
    Z2.test_gpu_vs_reference_on_z2(gtn_host_trunc, "block_trg_chi32", "block", "trg", 32, 3)
    assert _ops.SVD_PATH_STATS["truncated"] > before["truncated"]


# ---- flavour HOTRG (hotrg3dz: 6-leg tensors, bosonic legs, eig, hconjugate) on the Z2 tensor against the reference
@pytest.mark.parametrize("cut", [8, 16])
def test_host_hotrg3dz_z2_loose(gtn_host_trunc, cut):
    Z2.test_gpu_hotrg3dz_z2_loose(gtn_host_trunc, cut)


@pytest.mark.skipif(not __import__("os").environ.get("GTN_SLOW_TESTS"),
                    reason="~2 min of numpy SVDs of 4096 x 4096 matrices; set GTN_SLOW_TESTS=1")
def test_host_hotrg3dz_z2_chi64_vs_reference(gtn_host_trunc):
    """BASELINE.json configs[2] ('HOTRG chi=64'): Tnorm and F equal to the real reference to 1e-10"""
    Z2.test_gpu_hotrg3dz_z2_chi64_vs_reference(gtn_host_trunc)


# ---- large-dimension sweep / chains (tests/test_chains.py) through the host double
import test_chains as CH  # noqa: E402


@pytest.mark.parametrize("case", [0, 3, 4, 8, 10, 11])
def test_einsum_sweep_large_dims(gtn_host, case):
    CH.test_gpu_einsum_sweep_large_dims_vs_oracle(gtn_host, case)


def test_dense_data_inplace_edit_is_seen(gtn_host):
    GE.test_dense_data_inplace_edit_is_seen(gtn_host)


@pytest.mark.parametrize("fmt", ["dense", "block"])
@pytest.mark.parametrize("cplx", [True, False])
def test_unwritten_permutations_equal_written_ones(gtn_host, fmt, cplx):
    GE.test_unwritten_permutations_equal_written_ones(gtn_host, fmt, cplx)


@pytest.mark.parametrize("algo", ["trg", "atrg", "hotrg3dz"])
def test_steps_with_unwritten_permutations_equal_written_ones(gtn_host, algo):
    GE.test_steps_with_unwritten_permutations_equal_written_ones(gtn_host, algo)


@pytest.mark.parametrize("seed", range(12))
def test_random_chains_of_unwritten_permutations(gtn_host, seed):
    """random chains of signed permutations and conjugates (2-4 links) ending in a consumer that packs from the
    stored source -- a contraction with a partner, a trace, a decomposition -- against the same chain with every link
    written: identical bits (the composed GF(2) sign programs, incl. the p*p = p folding of traced pairs)"""
    import itertools
    import gtn_oracle as O
    from grassmanntn_b200 import _ops
    gtn = gtn_host
    rng = np.random.RandomState(1000 + seed)
    nleg = int(rng.choice([4, 4, 6]))
    dims = [int(rng.choice([2, 4])) for _ in range(nleg // 2)]
    shape = tuple(dims + dims)                       # legs a and a + nleg/2 have the same dimension ...
    stats = tuple([1] * (nleg // 2) + [-1] * (nleg // 2))       # ... and opposite statistics (traceable pairs)
    o = O.random_dense(shape, stats, dtype=complex, rng=rng)
    letters = "abcdef"[:nleg]
    links = []
    for _ in range(int(rng.randint(2, 5))):
        if rng.rand() < 0.3:
            links.append(("h", int(rng.randint(1, nleg))))
        else:
            links.append(("p", "".join(rng.permutation(list(letters)))))
    consumer = int(rng.randint(0, 3))

    def run(fmt):
        X = gtn.dense(o.data, statistics=stats)
        X = X.toblock() if fmt else X
        Y = X
        for kind, arg in links:
            if kind == "h":
                Y = Y.hconjugate(letters[:arg] + "|" + letters[arg:])
            else:
                Y = gtn.einsum(letters + "->" + arg, Y)
        outs = [Y]
        if consumer == 0:                            # contraction with its own conjugate over all but one leg
            Z = Y.hconjugate(letters[:1] + "|" + letters[1:])
            outs.append(gtn.einsum(letters + "," + letters[1:] + "z->" + letters[0] + "z", Y, Z))
        elif consumer == 1:                          # decomposition
            outs.extend(Y.svd(letters[:nleg // 2] + "|" + letters[nleg // 2:]))
        else:                                        # norm and scaled copy
            outs.append(Y * 0.25)
        return GE._bits(gtn, outs)
    saved = _ops.LAZY_PERMUTE
    try:
        for fmt in (False, True):
            _ops.LAZY_PERMUTE = False
            a = run(fmt)
            _ops.LAZY_PERMUTE = True
            b = run(fmt)
            assert GE._same(a, b), (links, consumer, fmt)
    finally:
        _ops.LAZY_PERMUTE = saved


def test_unwritten_permutation_is_not_an_alias(gtn_host):
    GE.test_unwritten_permutation_is_not_an_alias(gtn_host)


def test_perturbed_z2_chain_truncated_vs_reference(gtn_host_trunc):
    """the 12-step dcut-16 TRG chain of tests/test_chains.py (real-reference golden) with the subspace-iteration SVD
    and its certificate running on the host double: Tnorm and F to 1e-10 at every step (no graphs here; the GPU test
    adds the recorded step graph)"""
    import math
    import os
    gtn = gtn_host_trunc
    z = np.load(os.path.join(CH.G, "chains.npz"))
    ref = z["stepgraph_chain"]
    g = gtn.gauge2d
    T = gtn.dense(z["stepgraph_input"], statistics=tuple(int(s) for s in z["stepgraph_stats"])).toblock()
    logNorm = 0.0
    for i in range(12):
        T, Tn = g.trg(T, 16)
        logNorm = 2 * logNorm + math.log(Tn)
        F = (g.logZ(T, CH.BC) + logNorm) / 2 ** (i + 1)
        assert abs(Tn - ref[i, 0]) <= 1e-10 * ref[i, 0], (i, Tn, ref[i, 0])
        assert abs(F - complex(ref[i, 1], ref[i, 2])) <= 1e-10 * abs(F), (i, F)
    from grassmanntn_b200 import _ops
    assert _ops.SVD_PATH_STATS["truncated"] >= 8


def test_rank_deficient_sectors_chi64_rank_certificate(gtn_host_trunc):
    """TRG at chi = 64 on the Z2 tensor, first two steps against the oracle port (make_chain_goldens.py chi64).  The
    sector matrices of the second step (512 x 512) have exact rank 16 < k = 32: the number of values that pass the
    reference's rank rule must not be read off the noisy Ritz values -- the truncated path has to certify it through
    the deflated matrix (_engine.deflated_norm_bound) and come out with the reference's bond dimensions (32, 32)."""
    import math
    import os
    from grassmanntn_b200 import _engine as E, _ops
    gtn = gtn_host_trunc
    ref = np.load(os.path.join(CH.G, "chains.npz"))["oracle_trg_chi64"]
    g = gtn.gauge2d
    T = g.zcap(g.load_initial_tensor().toblock())
    before, calls = dict(_ops.SVD_PATH_STATS), dict(E.RANK_CHECK_STATS)
    logNorm = 0.0
    for i in range(2):
        T, Tn = g.trg(T, 64)
        logNorm = 2 * logNorm + math.log(Tn)
        F = (g.logZ(T, CH.BC) + logNorm) / 2 ** (i + 1)
        assert tuple(T.effective_shape[:2]) == (int(ref[i, 3]), int(ref[i, 4])), (i, T.effective_shape)
        assert abs(Tn - ref[i, 0]) <= 1e-10 * ref[i, 0], (i, Tn, ref[i, 0])
        assert abs(F - complex(ref[i, 1], ref[i, 2])) <= 1e-10 * abs(F), (i, F)
    assert _ops.SVD_PATH_STATS["truncated"] > before["truncated"]
    assert E.RANK_CHECK_STATS["certified"] > calls["certified"]


def test_truncated_eig_vs_oracle_D16(gtn_host_trunc):
    GE.test_truncated_eig_vs_oracle_D16(gtn_host_trunc)


# ---- initial-tensor compression pipeline (tests/test_tensor_prep.py) through the host double
import test_tensor_prep as TP  # noqa: E402


def test_tensor_preparation_vs_reference(gtn_host_trunc):
    TP.run_pipeline(gtn_host_trunc)


def test_bosonic_diagonal_einsum(gtn_host):
    TP.test_gpu_bosonic_diagonal_einsum(gtn_host)
