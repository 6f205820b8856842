"""GPU parity of the TMA-staged convolution kernel (csrc/conv_tma.cu) at launch shapes that exercise every part of its
tile plan: persistent multi-tile loops, halo slabs shared by the taps of one dimension, ragged tile edges (TMA
zero-fill on load, clipping on store), resident and streamed weight tiles, several N tiles, channel slices of wider
buffers, gradient accumulation through the bulk add-reduction, and the frame-parity slab types of the temporally
strided stem conv (forward and data gradient).  Truth: torch.nn.functional.conv3d in float64 -- the op the reference
reaches through nn.Conv3d (backbone/s3dg.py:11-13,39-42); the cp.async gather kernel is run on the same inputs as a
second opinion.  Tolerances as in test_conv_gpu.py."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from test_conv_gpu import _planes_of, _rel, _dump

pytestmark = pytest.mark.gpu

# (name, Cin, Cout, k, s, p, B, T, H, W)
CASES = [
    ("sp_64_192_big", 64, 192, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 3, 32, 32),
    ("sp_128_128_raggedH", 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1), 1, 2, 20, 16),
    ("sp_64_64_w8", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 1, 2, 16, 8),
    ("tm_192_192", 192, 192, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 16, 8, 8),
    ("tm_64_64_raggedT", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 12, 4, 8),
    ("stem_tm_resident", 64, 64, (7, 1, 1), (2, 1, 1), (3, 0, 0), 2, 32, 32, 32),
    ("pw_256_384", 256, 384, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 4, 16, 16),
    ("pw_480_192_partialK", 480, 192, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 4, 8, 8),
    ("pw_64_64_resident", 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 4, 8, 32, 32),
]
IDS = [c[0] for c in CASES]


def _geom(case):
    from coclr_b200 import ops
    name, Cin, Cout, k, s, p, B, T, H, W = case
    return ops, ops.Geometry(k, s, p)


def _takes_tma(cv):
    from coclr_b200 import lib as L
    info = (C.c_int * 8)()
    ok = L.load().coclr_conv_tma_plan(C.byref(cv), info)
    return bool(ok), list(info)


def _run(ops, cv, tma):
    from coclr_b200 import lib as L
    lib = L.load()
    lib.coclr_set_conv_tma(1 if tma else 0)
    try:
        L.check(lib.coclr_conv_igemm(C.byref(cv), L.num_sms(), L.stream_ptr()), "coclr_conv_igemm")
        torch.cuda.synchronize()
    finally:
        lib.coclr_set_conv_tma(1)


@pytest.mark.parametrize("case", CASES, ids=IDS)
@pytest.mark.parametrize("npass,bf16", [(3, 0), (1, 1)])
def test_tma_forward(case, npass, bf16, diag):
    ops, geom = _geom(case)
    name, Cin, Cout, k, s, p, B, T, H, W = case
    To, Ho, Wo = geom.out_dims(T, H, W)
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.relu(torch.randn(B, Cin, T, H, W, device="cuda", generator=g))
    w = torch.randn(Cout, Cin, *k, device="cuda", generator=g) * 0.05
    ref = F.conv3d(x.double(), w.double(), stride=s, padding=p)
    pl, Cp, coff = _planes_of(ops, x, bf16, lo=(npass > 1))
    pw = ops.PackedWeights(Cout, Cin, geom.taps, Cp, 0, bf16, "cuda").pack(w.contiguous())
    outs = {}
    for tma in (1, 0):
        dst = torch.full((B, To, Ho, Wo, Cout + 40), -3.0, device="cuda")
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
        cv = ops.make_conv(pl.src(coff, Cp, T, H, W), bf16, geom.c(0), B, (To, Ho, Wo), pw, dst, 32,
                           stats_sum=stats[:Cout], stats_sq=stats[Cout:], npass=npass)
        if tma:
            ok, info = _takes_tma(cv)
            diag["tma_plan/%s/p%d" % (name, npass)] = info
            assert ok, "the planner sends %s to the gather kernel" % name
        _run(ops, cv, tma)
        outs[tma] = (dst, stats)
    dst, stats = outs[1]
    got = dst[..., 32:32 + Cout].permute(0, 4, 1, 2, 3)
    err = _rel(got, ref)
    err_gather = _rel(outs[0][0][..., 32:32 + Cout].permute(0, 4, 1, 2, 3), ref)
    diag["tma_fwd/%s/p%d_bf%d" % (name, npass, bf16)] = [err, err_gather]
    tol = 2e-5 if npass == 3 and not bf16 else 3e-2
    if not err < tol:
        _dump("tma_fwd_%s_p%d" % (name, npass), got, ref)
    assert err < tol, "TMA forward %s rel err %.3e (gather kernel: %.3e)" % (name, err, err_gather)
    assert float((dst[..., :32] + 3.0).abs().max()) == 0.0          # neighbouring channels untouched
    assert float((dst[..., 32 + Cout:] + 3.0).abs().max()) == 0.0
    s1 = got.double().sum(dim=(0, 2, 3, 4))
    s2 = (got.double() ** 2).sum(dim=(0, 2, 3, 4))
    e1 = float((stats[:Cout] - s1).abs().max() / s1.abs().max().clamp_min(1e-30))
    e2 = float((stats[Cout:] - s2).abs().max() / s2.abs().max().clamp_min(1e-30))
    diag["tma_fwd_stats/%s/p%d_bf%d" % (name, npass, bf16)] = max(e1, e2)
    assert e1 < 1e-5 and e2 < 1e-5


@pytest.mark.parametrize("case", CASES, ids=IDS)
@pytest.mark.parametrize("accumulate", [0, 1], ids=["store", "reduce_add"])
def test_tma_dgrad(case, accumulate, diag):
    """Data gradient (transposed geometry) incl. the frame-parity tiles of the strided stem conv; accumulate = the
    bulk-tensor add-reduction into an existing gradient buffer (a tensor with several consumers)."""
    ops, geom = _geom(case)
    name, Cin, Cout, k, s, p, B, T, H, W = case
    To, Ho, Wo = geom.out_dims(T, H, W)
    g = torch.Generator(device="cuda").manual_seed(37)
    x = torch.randn(B, Cin, T, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, *k, device="cuda", generator=g) * 0.05
    dy = torch.randn(B, Cout, To, Ho, Wo, device="cuda", generator=g)
    xd = x.double().requires_grad_(True)
    (ref,) = torch.autograd.grad(F.conv3d(xd, w.double(), stride=s, padding=p), xd, dy.double())
    pl, Cop, coff = _planes_of(ops, dy, 1)
    pw = ops.PackedWeights(Cout, Cin, geom.taps, Cop, 1, 1, "cuda").pack(w.contiguous())
    base = torch.randn(B, T, H, W, Cin + 8, device="cuda", generator=g)
    res = {}
    for tma in (1, 0):
        dst = base.clone()
        cv = ops.make_conv(pl.src(coff, Cop, To, Ho, Wo), 1, geom.c(1), B, (T, H, W), pw, dst, 8,
                           accumulate=bool(accumulate), npass=3)
        if tma:
            ok, info = _takes_tma(cv)
            assert ok, "the planner sends the dgrad of %s to the gather kernel" % name
        _run(ops, cv, tma)
        out = dst[..., 8:] - (base[..., 8:] if accumulate else 0.0)
        res[tma] = _rel(out.permute(0, 4, 1, 2, 3), ref)
        assert torch.equal(dst[..., :8], base[..., :8])
        if tma and not res[tma] < 2e-4:
            _dump("tma_dgrad_%s_a%d" % (name, accumulate), out.permute(0, 4, 1, 2, 3), ref)
    diag["tma_dgrad/%s/acc%d" % (name, accumulate)] = [res[1], res[0]]
    assert res[1] < 2e-4, "TMA dgrad %s rel err %.3e (gather kernel: %.3e)" % (name, res[1], res[0])


def test_tma_matches_gather_bitwise_statistics_order_free(diag):
    """Same inputs, TMA kernel twice: outputs are bit-identical run to run (no atomics on the output path)."""
    ops, geom = _geom(CASES[0])
    name, Cin, Cout, k, s, p, B, T, H, W = CASES[0]
    g = torch.Generator(device="cuda").manual_seed(41)
    x = torch.relu(torch.randn(B, Cin, T, H, W, device="cuda", generator=g))
    w = torch.randn(Cout, Cin, *k, device="cuda", generator=g) * 0.05
    pl, Cp, coff = _planes_of(ops, x, 0)
    pw = ops.PackedWeights(Cout, Cin, geom.taps, Cp, 0, 0, "cuda").pack(w.contiguous())
    outs = []
    for _ in range(2):
        dst = torch.empty(B, T, H, W, Cout, device="cuda")
        cv = ops.make_conv(pl.src(coff, Cp, T, H, W), 0, geom.c(0), B, (T, H, W), pw, dst, 0, npass=3)
        _run(ops, cv, 1)
        outs.append(dst)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("npass,bf16", [(3, 0), (1, 1)])
def test_tma_window_stem(npass, bf16, diag):
    """The space-to-depth stem (backbone/s3dg.py:145 Conv_1a.conv1 as a stride-1 (1,4,4) conv over 16-channel pixels,
    engine.py Graph.stem_s2d): the four taps of a kernel row are ONE 128-byte run of the zero-padded input rows, read
    through a tensor map with overlapping rows; all 16 taps are served by one slab per tile and the weights stay
    resident.  Same conv on the gather kernel (padded source, pw = 0) as a second opinion."""
    from coclr_b200 import ops
    B, T, H2, W2, Cout = 4, 6, 32, 32, 64        # 192 tiles: more than one per SM, so the weights stay resident
    g = torch.Generator(device="cuda").manual_seed(43)
    x = torch.randn(B, 16, T, H2, W2, device="cuda", generator=g)
    x[:, 12:] = 0.0                                           # 4*Cin = 12 real channels
    w = torch.randn(Cout, 16, 1, 4, 4, device="cuda", generator=g) * 0.05
    ref = F.conv3d(x.double(), w.double(), padding=(0, 2, 2))[..., :H2, :W2]
    xpad = F.pad(x, (2, 2))                                   # the layout coclr_pack_input_s2d(pad_x=2) writes
    pl, Cp, coff = _planes_of(ops, xpad, bf16, ld_extra=0, coff=0, lo=(npass > 1))
    geom = ops.Geometry((1, 4, 4), (1, 1, 1), (0, 2, 0))
    pw = ops.PackedWeights(Cout, 16, 16, 16, 0, bf16, "cuda").pack(w.contiguous())
    errs = {}
    for tma in (1, 0):
        dst = torch.full((B, T, H2, W2, Cout), -3.0, device="cuda")
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
        cv = ops.make_conv(pl.src(0, 16, T, H2, W2 + 4), bf16, geom.c(0), B, (T, H2, W2), pw, dst, 0,
                           stats_sum=stats[:Cout], stats_sq=stats[Cout:], npass=npass)
        if tma:
            ok, info = _takes_tma(cv)
            diag["tma_plan/window_stem/p%d" % npass] = info
            assert ok and info[2] == 1, "window stem: planner %s" % info      # TMA kernel, resident weights
        _run(ops, cv, tma)
        got = dst.permute(0, 4, 1, 2, 3)
        errs[tma] = _rel(got, ref)
        if tma:
            s1 = got.double().sum(dim=(0, 2, 3, 4))
            assert float((stats[:Cout] - s1).abs().max() / s1.abs().max()) < 1e-5
    diag["tma_fwd/window_stem/p%d" % npass] = [errs[1], errs[0]]
    tol = 2e-5 if npass == 3 and not bf16 else 3e-2
    assert errs[0] < tol, "gather kernel on the padded stem input: %.3e" % errs[0]
    assert errs[1] < tol, "TMA window stem: %.3e (gather: %.3e)" % (errs[1], errs[0])
