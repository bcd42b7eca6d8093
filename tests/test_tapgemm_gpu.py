"""GPU parity of the tap-GEMM kernels (CUDA-core fp32, tcgen05 bf16, weight gradient) against the
plain-torch emulation in tests/emu.py, through the C-ABI."""
import pytest
import torch

from tests.emu import EmuOps, swizzle_slabs
from vsr_b200 import _lib as L
from vsr_b200.ops import TapTable

pytestmark = pytest.mark.gpu


def _ops():
    from vsr_b200.ops import cuda_ops
    return cuda_ops()


def conv3x3_table(kc, nt, n_k=1):
    taps = [(0, dy, dx, kb * kc) for dy in (-1, 0, 1) for dx in (-1, 0, 1) for kb in range(n_k)]
    return TapTable(kc=kc, nt=nt, groups=[(0, taps)])


def multi_src_table(kc, nt, n_src):
    return TapTable(kc=kc, nt=nt, groups=[(0, [(s, 0, 0, 0) for s in range(n_src)])])


def grouped_table(kc, nt, n_groups, n_phase):
    """deconv-like: each group has 4 shifted taps, writes its own channel slice; plus c0 offsets"""
    groups = []
    for g in range(n_groups):
        gy, gx = g // 2, g % 2
        taps = [(0, dy - 1 + gy, dx - 1 + gx, ((dy * 2 + dx) % n_phase) * kc) for dy in (0, 1) for dx in (0, 1)]
        groups.append((g * nt, taps))
    return TapTable(kc=kc, nt=nt, groups=groups)


def _setenv(monkeypatch, key, value):
    """tuning overrides are read once by the library: set the variable and have it re-read"""
    monkeypatch.setenv(key, value)
    _ops().lib.vsr_reload_tunables()


@pytest.fixture(autouse=True)
def _fresh_tunables():
    yield
    import os
    for k in [k for k in os.environ if k.startswith(("VSR_TC_", "VSR_WG_", "VSR_PDL"))]:
        del os.environ[k]
    _ops().lib.vsr_reload_tunables()


def _rand(shape, dtype, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(dtype)


def _run_case(tab, n, h, w, src_c, out_c, dtype, epi, seed=0, n_srcs=1):
    ops, emu = _ops(), EmuOps()
    gen = torch.Generator(device="cuda").manual_seed(seed)
    srcs = [_rand((n, h, w, src_c), dtype, gen) for _ in range(n_srcs)]
    w_plain = _rand((tab.n_taps_total, tab.nt, tab.kc), dtype, gen, scale=(tab.kc * max(1, tab.n_taps_total // tab.n_groups)) ** -0.5)
    w_dev = swizzle_slabs(w_plain) if dtype == torch.bfloat16 else w_plain.reshape(-1)
    bias = _rand((out_c,), torch.float32, gen)
    slope = torch.tensor([0.2], device="cuda")
    residual = _rand((n, h, w, out_c), dtype, gen)
    aux_y = _rand((n, h, w, out_c), dtype, gen)
    res2 = _rand((n, h, w, out_c), dtype, gen)
    outs = []
    for o in (ops, emu):
        out = torch.zeros((n, h, w, out_c), dtype=dtype, device="cuda")
        out2 = torch.zeros_like(out)
        part = torch.zeros(ops.partials_len, device="cuda")
        o.tapgemm(tab, srcs, out, w_dev, bias=bias, epi=epi, out_scale=0.5, slope=slope, residual=residual,
                  aux_y=aux_y, out2=out2, res2=res2, slope_partials=part)
        outs.append((out.float(), out2.float(), part.sum()))
    torch.cuda.synchronize()
    (a, a2, ap), (b, b2, bp) = outs
    # bf16: kernel and emulation accumulate the same bf16 inputs in fp32; after the final rounding to bf16 they can differ
    # by one ulp where the fp32 sums straddle a rounding boundary (<= 2^-7 of the tensor maximum): 8.5e-3 leaves no room
    # for a dropped tap (one of 64 taps is ~3 % of the output range)
    tol = 8.5e-3 if dtype == torch.bfloat16 else 2e-5
    scale = b.abs().max().item() + 1e-6
    assert (a - b).abs().max().item() / scale < tol, f"out mismatch {(a - b).abs().max().item() / scale}"
    if epi & L.EPI_OUT2:
        s2 = b2.abs().max().item() + 1e-6
        assert (a2 - b2).abs().max().item() / s2 < tol
    if epi & L.EPI_PRELU_BWD:
        assert abs(ap.item() - bp.item()) <= 2e-2 * max(1.0, abs(bp.item())) if dtype == torch.bfloat16 \
            else abs(ap.item() - bp.item()) <= 1e-3 * max(1.0, abs(bp.item()))


EPIS = [0, L.EPI_BIAS | L.EPI_PRELU, L.EPI_BIAS | L.EPI_SCALE | L.EPI_RES_PRE | L.EPI_RELU,
        L.EPI_RES_PRE | L.EPI_PRELU_BWD, L.EPI_BIAS | L.EPI_PRELU | L.EPI_OUT2, L.EPI_RELU_BWD]


@pytest.mark.parametrize("epi", EPIS)
def test_simt_fp32_conv3x3(epi):
    _run_case(conv3x3_table(32, 48), n=2, h=9, w=13, src_c=32, out_c=48, dtype=torch.float32, epi=epi)


def test_simt_fp32_multi_src_ragged():
    _run_case(multi_src_table(24, 40, 3), n=3, h=5, w=7, src_c=24, out_c=40, dtype=torch.float32,
              epi=L.EPI_BIAS | L.EPI_PRELU, n_srcs=3)


def test_simt_fp32_grouped():
    _run_case(grouped_table(16, 32, 4, 3), n=2, h=6, w=6, src_c=48, out_c=128, dtype=torch.float32, epi=L.EPI_BIAS)


def test_tc_bf16_1x1_single_tile():
    # smallest possible tcgen05 problem: one tap, one tile
    _run_case(multi_src_table(64, 64, 1), n=1, h=4, w=32, src_c=64, out_c=64, dtype=torch.bfloat16, epi=0)


def test_tc_bf16_1x1_k_advance():
    # 4 sources -> 4 taps: exercises the stage ring and accumulation
    _run_case(multi_src_table(64, 64, 4), n=2, h=8, w=32, src_c=64, out_c=64, dtype=torch.bfloat16, epi=0, n_srcs=4)


@pytest.mark.parametrize("epi", EPIS)
def test_tc_bf16_conv3x3_epilogues(epi):
    _run_case(conv3x3_table(64, 64), n=2, h=32, w=32, src_c=64, out_c=64, dtype=torch.bfloat16, epi=epi)


def test_tc_bf16_conv3x3_n256():
    _run_case(conv3x3_table(64, 256), n=2, h=32, w=32, src_c=64, out_c=256, dtype=torch.bfloat16,
              epi=L.EPI_BIAS)


def test_tc_bf16_ragged_tiles():
    # H, W not multiples of the pixel box: partial tiles + TMA zero fill on every side
    _run_case(conv3x3_table(64, 128), n=3, h=19, w=21, src_c=64, out_c=128, dtype=torch.bfloat16,
              epi=L.EPI_BIAS | L.EPI_PRELU)


def test_tc_bf16_grouped_channel_slices():
    _run_case(grouped_table(64, 256, 4, 3), n=2, h=16, w=16, src_c=192, out_c=1024, dtype=torch.bfloat16,
              epi=L.EPI_BIAS | L.EPI_PRELU)


def test_tc_bf16_many_taps_many_tiles():
    # 64 taps (strided-conv shape), more tiles than SMs -> double-buffered TMEM, phase wrap
    taps = [(0, (t // 8) % 3 - 1, (t % 8) % 3 - 1, (t % 16) * 64) for t in range(64)]
    tab = TapTable(kc=64, nt=64, groups=[(0, taps)])
    _run_case(tab, n=8, h=32, w=32, src_c=1024, out_c=64, dtype=torch.bfloat16, epi=L.EPI_BIAS | L.EPI_PRELU)


def test_tc_matches_simt_bf16():
    ops = _ops()
    gen = torch.Generator(device="cuda").manual_seed(3)
    tab = conv3x3_table(64, 64)
    src = _rand((2, 16, 16, 64), torch.bfloat16, gen)
    w_plain = _rand((9, 64, 64), torch.bfloat16, gen, 0.05)
    o1 = torch.zeros((2, 16, 16, 64), dtype=torch.bfloat16, device="cuda")
    o2 = torch.zeros_like(o1)
    ops.tapgemm(tab, [src], o1, swizzle_slabs(w_plain))
    ops.tapgemm(tab, [src], o2, w_plain.reshape(-1), force_simt=True)
    torch.cuda.synchronize()
    assert (o1.float() - o2.float()).abs().max().item() <= 2e-2 * o2.float().abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_wgrad(dtype):
    ops, emu = _ops(), EmuOps()
    gen = torch.Generator(device="cuda").manual_seed(5)
    kc = 64 if dtype == torch.bfloat16 else 24
    tab = grouped_table(kc, 32, 2, 2)
    n, h, w = 3, 11, 9
    src = _rand((n, h, w, 2 * kc), dtype, gen)
    dz = _rand((n, h, w, 64), dtype, gen)
    res = []
    for o in (ops, emu):
        dw = torch.ones(tab.n_taps_total * tab.nt * tab.kc, device="cuda")
        ws = torch.empty(max(16, o.tapgemm_wgrad_workspace(tab, [src], dz)) // 4, device="cuda")
        o.tapgemm_wgrad(tab, [src], dz, dw, True, ws)
        res.append(dw)
    torch.cuda.synchronize()
    assert (res[0] - res[1]).abs().max().item() <= 1e-3 * res[1].abs().max().item()
    # determinism: a second run is bit-identical
    dw2 = torch.ones_like(res[0])
    ws = torch.empty(max(16, ops.tapgemm_wgrad_workspace(tab, [src], dz)) // 4, device="cuda")
    ops.tapgemm_wgrad(tab, [src], dz, dw2, True, ws)
    torch.cuda.synchronize()
    assert torch.equal(dw2, res[0])


def _wgrad_case(tab, n, h, w, src_c, out_c, n_srcs=1, seed=11):
    ops, emu = _ops(), EmuOps()
    gen = torch.Generator(device="cuda").manual_seed(seed)
    srcs = [_rand((n, h, w, src_c), torch.bfloat16, gen) for _ in range(n_srcs)]
    dz = _rand((n, h, w, out_c), torch.bfloat16, gen)
    res = []
    for o in (ops, emu):
        dw = torch.ones(tab.n_taps_total * tab.nt * tab.kc, device="cuda")
        ws = torch.empty(max(16, o.tapgemm_wgrad_workspace(tab, srcs, dz)) // 4 + 4, device="cuda")
        o.tapgemm_wgrad(tab, srcs, dz, dw, True, ws)
        res.append(dw)
    torch.cuda.synchronize()
    err = (res[0] - res[1]).abs().max().item() / res[1].abs().max().item()
    assert err <= 1e-3, err
    dw2 = torch.ones_like(res[0])
    ws = torch.empty(max(16, ops.tapgemm_wgrad_workspace(tab, srcs, dz)) // 4 + 4, device="cuda")
    ops.tapgemm_wgrad(tab, srcs, dz, dw2, True, ws)
    torch.cuda.synchronize()
    assert torch.equal(dw2, res[0])


def test_wgrad_tc_single_tap_single_tile():
    _wgrad_case(multi_src_table(64, 64, 1), n=1, h=4, w=32, src_c=64, out_c=64)


def test_wgrad_tc_1x1_multi_src_many_tiles():
    _wgrad_case(multi_src_table(64, 64, 5), n=4, h=32, w=64, src_c=64, out_c=64, n_srcs=5)


def test_wgrad_tc_conv3x3_n256_ragged():
    _wgrad_case(conv3x3_table(64, 256), n=3, h=19, w=21, src_c=64, out_c=256)


def test_wgrad_tc_strided_64taps():
    taps = [(0, (t // 8) % 3 - 1, (t % 8) % 3 - 1, (t % 16) * 64) for t in range(64)]
    _wgrad_case(TapTable(kc=64, nt=64, groups=[(0, taps)]), n=4, h=16, w=16, src_c=1024, out_c=64)


def test_wgrad_tc_grouped_nt128_nonuniform_groups():
    g0 = [(0, dy, dx, 0) for dy in (0, 1) for dx in (0, 1)]
    g1 = [(0, dy, dx, 64) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    tab = TapTable(kc=64, nt=128, groups=[(0, g0), (128, g1)])
    _wgrad_case(tab, n=2, h=12, w=20, src_c=128, out_c=256)


def strided_conv_table(r=4, k=8, p=2, F=64):
    """nn.Conv2d(k, stride=r, pad=p) on a phase-blocked r-times map (row-major slots): k*k taps in one group."""
    taps = []
    for ky in range(k):
        for kx in range(k):
            dy, py = divmod(ky - p, r)
            dx, px = divmod(kx - p, r)
            taps.append((0, dy, dx, (py * r + px) * F))
    return TapTable(kc=F, nt=F, groups=[(0, taps)])


@pytest.mark.parametrize("h,w", [(32, 32), (19, 32), (12, 12)])
@pytest.mark.parametrize("epi", [L.EPI_BIAS | L.EPI_PRELU, L.EPI_PRELU_BWD, L.EPI_RES_PRE | L.EPI_PRELU_BWD])
def test_tc_bf16_strided_conv_shared_loads(h, w, epi, monkeypatch):
    """taps that differ by a row shift share one A box and two stacked pixel tiles share the slabs
    (tapgemm_tc2.cu shared-load mode); the same launch with the mode off must agree with the emulation too."""
    tab = strided_conv_table()
    _run_case(tab, n=3, h=h, w=w, src_c=1024, out_c=64, dtype=torch.bfloat16, epi=epi, seed=21)
    _setenv(monkeypatch, "VSR_TC_TALL", "0")
    _run_case(tab, n=3, h=h, w=w, src_c=1024, out_c=64, dtype=torch.bfloat16, epi=epi, seed=21)


def test_tc_bf16_strided_conv_r2_three_row_columns():
    # 6x6 stride 2 pad 2: every (slot, dx) column has three row shifts
    _run_case(strided_conv_table(r=2, k=6), n=2, h=24, w=16, src_c=256, out_c=64, dtype=torch.bfloat16,
              epi=L.EPI_BIAS | L.EPI_PRELU, seed=22)


def test_tc_bf16_resident_and_streamed_weights_agree(monkeypatch):
    # deconv shape: 4 groups x 4 taps, nt 256; enough tiles for the weight-resident mode
    tab = grouped_table(64, 256, 4, 1)
    for mode in ("1", "0"):
        _setenv(monkeypatch, "VSR_TC_RESIDENT", mode)
        _run_case(tab, n=20, h=32, w=32, src_c=64, out_c=1024, dtype=torch.bfloat16, epi=L.EPI_BIAS | L.EPI_PRELU, seed=23)


def test_tc_bf16_without_programmatic_dependent_launch(monkeypatch):
    tab = conv3x3_table(64, 64)
    _setenv(monkeypatch, "VSR_PDL", "0")
    _run_case(tab, n=2, h=20, w=32, src_c=64, out_c=64, dtype=torch.bfloat16, epi=L.EPI_BIAS | L.EPI_PRELU, seed=24)


def test_wgrad_tc_shared_pair_loads(monkeypatch):
    _setenv(monkeypatch, "VSR_WG_TALL", "1")
    _wgrad_case(strided_conv_table(), n=4, h=16, w=32, src_c=1024, out_c=64, seed=25)
    _wgrad_case(strided_conv_table(), n=2, h=19, w=32, src_c=1024, out_c=64, seed=26)


@pytest.mark.parametrize("n_groups,nt", [(4, 128), (2, 256), (4, 64)])
def test_tc_bf16_shared_loads_several_groups_many_tiles(n_groups, nt, monkeypatch):
    """deconv-like tables (every group: two columns of two row shifts) with more tiles than CTAs and a ring
    shallower than the producer count - the case that exposed the mbarrier parity aliasing of a producer
    that runs a whole ring ahead (tapgemm_tc2.cu: at most `stages` producers are active)."""
    groups = []
    for g in range(n_groups):
        gy, gx = g // 2, g % 2
        groups.append((g * nt, [(0, dy - 1 + gy, dx - 1 + gx, 0) for dy in (0, 1) for dx in (0, 1)]))
    tab = TapTable(kc=64, nt=nt, groups=groups)
    _setenv(monkeypatch, "VSR_TC_RESIDENT", "0")
    _run_case(tab, n=20, h=32, w=32, src_c=64, out_c=n_groups * nt, dtype=torch.bfloat16, epi=L.EPI_BIAS | L.EPI_PRELU, seed=27)
    _setenv(monkeypatch, "VSR_TC_STAGES", "2")
    _run_case(tab, n=20, h=32, w=32, src_c=64, out_c=n_groups * nt, dtype=torch.bfloat16, epi=L.EPI_PRELU_BWD, seed=28)
    _setenv(monkeypatch, "VSR_TC_STAGES", "1")
    _run_case(tab, n=6, h=32, w=32, src_c=64, out_c=n_groups * nt, dtype=torch.bfloat16, epi=0, seed=29)


@pytest.mark.parametrize("pair", ["0", "1"])
def test_tc_bf16_cta_pairs_and_single_ctas_agree_with_the_emulation(pair, monkeypatch):
    """every mode of the kernel with CTA pairs forced on (tcgen05 cta_group::2: M = 256 over two CTAs, half of every
    weight slab per CTA) and forced off: odd pixel-tile counts (the last pair's second CTA works on an out-of-range
    tile), resident weights with several groups, shared loads with two stacked sub-tiles, streamed slabs, every
    epilogue operand path"""
    _setenv(monkeypatch, "VSR_TC_PAIR", pair)
    # 3 pixel tiles (odd), one tap, plain epilogue
    _run_case(multi_src_table(64, 64, 1), n=3, h=4, w=32, src_c=64, out_c=64, dtype=torch.bfloat16, epi=0, seed=31)
    # 3x3, partial tiles on every side, odd tile count, bias + PReLU
    _run_case(conv3x3_table(64, 128), n=3, h=19, w=21, src_c=64, out_c=128, dtype=torch.bfloat16, epi=L.EPI_BIAS | L.EPI_PRELU, seed=32)
    # deconv shape: 4 groups x 4 taps, nt 256, resident weights, contiguous tile ranges
    _run_case(grouped_table(64, 256, 4, 1), n=21, h=32, w=32, src_c=64, out_c=1024, dtype=torch.bfloat16,
              epi=L.EPI_BIAS | L.EPI_PRELU, seed=33)
    # strided-convolution shape: 64 taps, nt 64, shared loads, PReLU' with residual
    _run_case(strided_conv_table(), n=5, h=32, w=32, src_c=1024, out_c=64, dtype=torch.bfloat16,
              epi=L.EPI_RES_PRE | L.EPI_PRELU_BWD, seed=34)
    # second output + second residual, several sources
    _run_case(multi_src_table(64, 64, 4), n=7, h=8, w=32, src_c=64, out_c=64, dtype=torch.bfloat16,
              epi=L.EPI_BIAS | L.EPI_PRELU | L.EPI_OUT2, seed=35, n_srcs=4)
    # more tiles than CTA pairs, two TMEM buffers in flight, PReLU' on a wide output
    _run_case(grouped_table(64, 256, 4, 1), n=40, h=32, w=32, src_c=64, out_c=1024, dtype=torch.bfloat16,
              epi=L.EPI_PRELU_BWD, seed=36)


@pytest.mark.parametrize("shape,ntaps", [((3, 16, 24), [2, 3]), ((2, 8, 512), [4, 5, 6]), ((5, 7, 9), [1]), ((1, 32, 32), [2, 3, 4, 5])])
def test_wgrad_shared_matches_float64(shape, ntaps):
    """csrc/wgrad_shared.cu: the weight (+ bias) gradients of several 1x1 convolutions over one list of 64-channel maps in
    one pass (every source tile loaded once) against float64 sums over the same bf16 operands; ragged pixel grids (partial
    TMA boxes), odd tap counts (the last pair's second half is padding), accumulate on / off"""
    from vsr_b200.ops import cuda_ops
    ops = cuda_ops()
    g = torch.Generator(device="cuda").manual_seed(sum(ntaps))
    n, h, w = shape
    srcs = [torch.randn(n, h, w, 64, device="cuda", generator=g).bfloat16() for _ in range(max(ntaps))]
    dzs = [torch.randn(n, h, w, 64, device="cuda", generator=g).bfloat16() for _ in ntaps]
    assert ops.wgrad_shared_ok(srcs, dzs, ntaps)
    dws = [torch.full((nt, 64, 64), 0.5, device="cuda") for nt in ntaps]
    dbs = [torch.full((64,), -1.0, device="cuda") if i % 2 == 0 else None for i in range(len(ntaps))]
    ws = {}

    def workspace_of(nbytes):
        ws["t"] = torch.empty((nbytes + 3) // 4, device="cuda")
        return ws["t"]

    for accumulate in (False, True):
        ops.wgrad_shared(srcs, dzs, ntaps, dws, dbs, accumulate, workspace_of)
        mult = 2.0 if accumulate else 1.0
        for dz, nt, dw, db in zip(dzs, ntaps, dws, dbs):
            z = dz.reshape(-1, 64).double()
            ref = torch.stack([z.t() @ srcs[t].reshape(-1, 64).double() for t in range(nt)])
            assert (dw.double() - mult * ref).abs().max() <= 2e-5 * mult * ref.abs().max()
            if db is not None:
                cs = z.sum(0)
                assert (db.double() - mult * cs).abs().max() <= 2e-5 * mult * cs.abs().max() + 1e-4
    assert not ops.wgrad_shared_ok(srcs + srcs, dzs, ntaps) or len(srcs) <= 4
