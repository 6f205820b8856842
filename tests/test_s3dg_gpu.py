"""GPU parity of the S3D-G variant (`--net s3dg`, reference backbone/select_backbone.py:8-9: S3D with a SelfGating,
backbone/s3dg.py:68-78, on every SepInception branch output, :107-112,125-129) through model.pretrain.InfoNCE -> C ABI
at a CPU-checkable shape (bs 4, 8 frames, 64x64, K=128), against
  (a) the golden vectors produced by the unmodified reference (tests/golden/infonce_s3dg.npz),
  (b) the oracle (oracle/coclr_oracle.py) evaluated in float64 on the GPU as ground truth,
plus the six gating entry points of include/coclr_b200.h one by one against float64 torch.
Tolerance on logits / loss / queue: 1e-3 relative (north star). Gradients: oracle-relative budget."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
pytestmark = pytest.mark.gpu

GOLD = os.path.join(ROOT, "tests", "golden", "infonce_s3dg.npz")


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------
# the kernels, one by one
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,P,members,lo", [(3, 37, [24, 40, 8, 16], True), (2, 512, [64, 128, 32, 32], True),
                                            (4, 64, [384, 384, 128, 128], True), (2, 50, [16, 8, 8, 8], False)],
                         ids=["ragged", "mixed3b", "mixed5c", "single_plane"])
def test_gating_kernels_vs_float64(B, P, members, lo, diag):
    """forward: planes -> mean -> gate -> planes scaled in place; backward: dgate, fc gradients, gradient w.r.t. the
    un-gated activation -- against autograd through the reference's formula in float64."""
    from coclr_b200 import lib as L, ops
    lib = L.load()
    dev = "cuda"
    Cc = sum(members)
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + P)
    y = torch.randn(B, P, Cc, generator=g).to(dev)                       # raw conv output
    scale = (torch.rand(Cc, generator=g) + 0.5).to(dev)
    shift = (torch.randn(Cc, generator=g) * 0.3).to(dev)
    a = torch.relu(y * scale + shift)
    pl = ops.Planes((B, 1, 1, P, Cc), 0, dev, lo=lo)
    ops.split_into(a.view(B * P, Cc).contiguous(), pl)
    a_pl = pl.value().double().view(B, P, Cc).clone()                    # what the planes hold (exact input)
    Ws = [(torch.randn(n, n, generator=g) * (4.0 / n) ** 0.5).to(dev) for n in members]
    bs = [(torch.randn(n, generator=g) * 0.2).to(dev) for n in members]
    mean = torch.empty(B, Cc, device=dev)
    gate = torch.empty(B, Cc, device=dev)
    s = L.stream_ptr()
    L.check(lib.coclr_gate_mean(L.dptr(pl.hi), L.dptr(pl.lo), pl.bf16, pl.ld, B, P, Cc, L.dptr(mean), s), "coclr_gate_mean")
    off = 0
    for W, b in zip(Ws, bs):
        L.check(lib.coclr_gate_fc(L.dptr(mean), L.dptr(W), L.dptr(b), L.dptr(gate), B, Cc, off, W.shape[0], s), "coclr_gate_fc")
        off += W.shape[0]
    L.check(lib.coclr_gate_apply(L.dptr(pl.hi), L.dptr(pl.lo), pl.bf16, pl.ld, B, P, Cc, L.dptr(gate), s), "coclr_gate_apply")
    # float64 reference, autograd through (scale, shift fixed): a -> out
    a64 = a_pl.clone().requires_grad_(True)
    W64 = [w.double().clone().requires_grad_(True) for w in Ws]
    b64 = [b.double().clone().requires_grad_(True) for b in bs]
    m64 = a64.mean(dim=1)
    outs, off = [], 0
    gates64 = []
    for W, b in zip(W64, b64):
        n = W.shape[0]
        gt = torch.sigmoid(torch.nn.functional.linear(m64[:, off:off + n], W, b))
        gates64.append(gt)
        outs.append(gt[:, None, :] * a64[:, :, off:off + n])
        off += n
    out64 = torch.cat(outs, dim=2)
    e_mean = _rel(mean, m64.detach())
    e_gate = _rel(gate, torch.cat(gates64, 1).detach())
    e_out = _rel(pl.value().view(B, P, Cc), out64.detach())
    assert e_mean < 2e-6 and e_gate < 2e-6, (e_mean, e_gate)
    assert e_out < (2e-6 if lo else 1e-3), e_out                          # a single fp16 plane holds 11 bits
    # backward.  The kernels recompute the activation from (y, scale, shift) in fp32; the reference differentiates at
    # the plane values, so compare with the plane-exact float32 activation fed to both (lo planes: equal to ~2^-22).
    dout = torch.randn(B, P, Cc, generator=g).to(dev)
    out64.backward(dout.double())
    dgate = torch.empty(B, Cc, device=dev)
    dmean = torch.empty(B, Cc, device=dev)
    dWs = [torch.empty_like(w) for w in Ws]
    dbs = [torch.empty_like(b) for b in bs]
    d_inplace = dout.clone()
    L.check(lib.coclr_gate_bwd_reduce(L.dptr(y), Cc, L.dptr(scale), L.dptr(shift), 1, L.dptr(dout), Cc, B, P, Cc,
                                      L.dptr(dgate), s), "coclr_gate_bwd_reduce")
    off = 0
    for W, dW, db in zip(Ws, dWs, dbs):
        n = W.shape[0]
        L.check(lib.coclr_gate_fc_bwd(L.dptr(dgate), L.dptr(gate), L.dptr(mean), L.dptr(W), L.dptr(dW), L.dptr(db),
                                      L.dptr(dmean), B, Cc, off, n, s), "coclr_gate_fc_bwd")
        off += n
    L.check(lib.coclr_gate_bwd_apply(L.dptr(d_inplace), Cc, L.dptr(gate), L.dptr(dmean), B, P, Cc, s), "coclr_gate_bwd_apply")
    torch.cuda.synchronize()
    tol = 5e-6 if lo else 2e-3
    errs = {"da": _rel_l2(d_inplace, a64.grad)}
    for i, (dW, db) in enumerate(zip(dWs, dbs)):
        errs["dW%d" % i] = _rel_l2(dW, W64[i].grad)
        errs["db%d" % i] = _rel_l2(db, b64[i].grad)
    diag["s3dg/kernels/B%d_P%d_C%d_lo%d" % (B, P, Cc, int(lo))] = dict(mean=e_mean, gate=e_gate, out=e_out, **errs)
    assert max(errs.values()) < tol, errs


# ---------------------------------------------------------------------------------------------
# the InfoNCE step
# ---------------------------------------------------------------------------------------------
def _build(precision="parity"):
    import make_golden as MG
    from model.pretrain import InfoNCE
    from oracle import coclr_oracle as O
    c = MG.S3DG_CFG
    torch.manual_seed(0)
    model = InfoNCE("s3dg", 128, c["K"], 0.999, 0.07, precision=precision)
    sd = O.synth_state(O.infonce_shapes(128, c["K"], network="s3dg"), seed=0, ptr=c["ptr"])
    model.load_state_dict(O.with_aliases(sd), strict=True)
    return model.cuda().train(), sd


def _oracle_truth(sd, block, idx, dtype, record=None):
    from oracle import coclr_oracle as O
    sdd = {}
    for k, v in sd.items():
        t = v.detach().clone().cuda()
        sdd[k] = t.to(dtype) if t.is_floating_point() else t
    for k in O.param_keys(sdd, "encoder_q."):
        sdd[k].requires_grad_(True)
    orig = O.sep_inception
    if record is not None:
        def rec(s, pre, x, training):
            y = orig(s, pre, x, training)
            record[pre] = y.detach()
            return y
        O.sep_inception = rec
    try:
        logits, labels = O.infonce_forward(sdd, [block.cuda().to(dtype)], idx)
    finally:
        O.sep_inception = orig
    loss = O.infonce_loss(logits[0], labels.cuda())
    loss.backward()
    return sdd, logits[0], loss


@pytest.fixture(scope="module")
def step():
    import make_golden as MG
    from coclr_b200 import moco
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = MG.S3DG_CFG
    model, sd = _build()
    block = MG.make_inputs_shifted(c["B"], c["T"], c["HW"]).cuda()
    torch.manual_seed(77)
    idx = torch.randperm(c["B"])
    torch.manual_seed(77)
    logits, labels = model(block)
    loss = moco.nce_cross_entropy(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    rec64 = {}
    truth = _oracle_truth(sd, block, idx, torch.float64, rec64)
    ref32 = _oracle_truth(sd, block, idx, torch.float32)
    return dict(model=model, sd=sd, logits=logits, labels=labels, loss=loss, truth=truth, ref32=ref32, rec64=rec64,
                block=block, cfg=c)


def test_s3dg_forward_matches_golden_and_truth(step, diag):
    gold = np.load(GOLD)
    logits, loss, model = step["logits"], step["loss"], step["model"]
    _, lg64, loss64 = step["truth"]
    e_gold, e_true = _rel(logits, gold["logits"]), _rel(logits, lg64)
    diag["s3dg/logits_vs_golden"] = e_gold
    diag["s3dg/logits_vs_fp64"] = e_true
    diag["s3dg/fp32oracle_vs_fp64"] = _rel(step["ref32"][1], lg64)
    diag["s3dg/loss"] = [float(loss), float(gold["loss"]), float(loss64)]
    assert logits.shape == (4, 129)
    assert e_gold < 1e-3 and e_true < 1e-3
    assert abs(float(loss) - float(gold["loss"])) < 1e-3 * max(1.0, abs(float(gold["loss"])))
    e_q = _rel(model.queue, gold["queue"])
    diag["s3dg/queue_vs_golden"] = e_q
    assert e_q < 1e-3
    assert int(model.queue_ptr) == int(gold["queue_ptr"][0]) == 20


def test_s3dg_ema_covers_gating_parameters(step):
    """_momentum_update_key_encoder runs over every parameter, the SelfGating Linear layers included
    (model/pretrain.py:76-80): bit-exact against the fp32 oracle and the reference golden."""
    gold = np.load(GOLD)
    msd = step["model"].state_dict()
    k = "encoder_k.0.Mixed_4b.gating_b2.fc.weight"
    assert _rel(msd[k], gold["ema/" + k]) < 1e-6
    assert torch.equal(msd[k].cpu(), step["ref32"][0][k].cpu())
    assert not torch.equal(msd[k].cpu(), step["sd"][k])


def test_s3dg_block_outputs(step, diag):
    """Gated output of each of the 9 SepInception blocks of encoder_q against the float64 oracle."""
    model, rec, c = step["model"], step["rec64"], step["cfg"]
    eng = model.encoder_q._engine
    plan = eng.plans[(c["B"], c["T"], c["HW"], c["HW"], True, True)]
    per = {}
    for t in eng.graph.tensors:
        if not t.gates:
            continue
        y = plan.acts[t.index].pl.value()
        ref = rec["encoder_q." + t.name].permute(0, 2, 3, 4, 1)
        per[t.name] = _rel_l2(y, ref)
    assert len(per) == 9
    diag["s3dg/blocks"] = per
    assert max(per.values()) < 1e-3, per


def test_s3dg_gradients_oracle_relative(step, diag):
    """Every parameter gradient of encoder_q against the float64 oracle, next to the float32 oracle's own error.
    What sets both errors is not arithmetic precision but ReLU decisions: an activation whose pre-ReLU value is within
    the forward error eps of zero takes the other branch, and k such flips among N active elements cost sqrt(k / N) in
    relative L2 -- about sqrt(0.8 * eps) per block whatever its size.  With eps = 2e-4 at Mixed_5c
    (test_s3dg_block_outputs; 64 BatchNorm samples per channel at this shape) that is ~1.3e-2 for the last block alone;
    measured here: median 3.0e-2 for this implementation, 0.9e-2 for float32 PyTorch (whose last block happened to see
    no flip: 1.4e-4 there), gating parameters no different from the convolutions around them.  Budget: 4x the float32
    oracle's median, 5e-2 (or 5x the float32 oracle) per tensor with at most two exceptions, 0.2 hard."""
    import make_golden as MG
    model = step["model"]
    sd64, sd32 = step["truth"][0], step["ref32"][0]
    named = dict(model.named_parameters())
    out, bad = {}, []
    for k in sorted(sd64):
        if not k.startswith("encoder_q.") or not (k.endswith(".weight") or k.endswith(".bias")):
            continue
        g64 = sd64[k].grad
        e_new, e_ref = _rel_l2(named[k].grad, g64), _rel_l2(sd32[k].grad, g64)
        out[k] = [e_new, e_ref]
        if not e_new < max(5 * e_ref, 5e-2):
            bad.append((k, e_new, e_ref))
        assert e_new < 0.2, (k, e_new, e_ref)
    diag["s3dg/grad_err_new_vs_ref"] = out
    med_new = float(np.median([v[0] for v in out.values()]))
    med_ref = float(np.median([v[1] for v in out.values()]))
    diag["s3dg/grad_median_new_ref"] = [med_new, med_ref]
    gate_errs = [v[0] for k, v in out.items() if ".gating_b" in k]
    diag["s3dg/grad_gating_max_median"] = [max(gate_errs), float(np.median(gate_errs))]
    assert len(gate_errs) == 72
    assert med_new < 4 * med_ref + 1e-3, (med_new, med_ref)
    gold = np.load(GOLD)
    vs_gold = {}
    for k in MG.GRAD_KEYS_S3DG:
        got, _ = MG.compact(named[k].grad.cpu().numpy())
        vs_gold[k] = _rel_l2(got, gold["grad/" + k])
    diag["s3dg/grad_vs_golden"] = vs_gold
    assert len(bad) <= 2, bad[:5]


def test_s3dg_adam_and_second_step(step, diag):
    """FlatAdam moves the gating parameters like the oracle's Adam on the same gradients; a second forward (CUDA-graph
    replay with the updated parameters) stays within 1e-3 of the float64 oracle started from the same state."""
    import make_golden as MG
    from coclr_b200 import moco
    from oracle import coclr_oracle as O
    model, c = step["model"], step["cfg"]
    enc = model.encoder_q
    named = dict(enc.named_parameters())
    params = {k: v.detach().clone() for k, v in named.items()}
    grads = {k: v.grad.detach().clone() for k, v in named.items()}
    O.adam_step(params, grads, {}, lr=1e-3, weight_decay=1e-5)
    opt = moco.FlatAdam(enc, lr=1e-3, weight_decay=1e-5)
    opt.step()
    torch.cuda.synchronize()
    worst = max(_rel(named[k], params[k]) for k in named)
    diag["s3dg/adam"] = worst
    assert worst < 1e-5
    assert sum(1 for k in named if ".gating_b" in k) == 72
    msd = model.state_dict()
    sd64 = {k: (msd[k].detach().clone().double() if msd[k].is_floating_point() else msd[k].detach().clone())
            for k in step["sd"]}
    block = MG.make_inputs_shifted(c["B"], c["T"], c["HW"], seed=99).cuda()
    torch.manual_seed(78)
    idx = torch.randperm(c["B"])
    torch.manual_seed(78)
    logits, labels = model(block)
    for k in O.param_keys(sd64, "encoder_q."):       # train mode is recognised from q.requires_grad (pretrain.py:157)
        sd64[k].requires_grad_(True)
    lg64, _ = O.infonce_forward(sd64, [block.double()], idx)
    e = _rel(logits, lg64[0].detach())
    diag["s3dg/second_step_logits_vs_fp64"] = e
    assert e < 1e-3
    assert _rel(model.queue, sd64["queue"]) < 1e-3


def test_s3dg_bare_backbone_forward(diag):
    """select_backbone('s3dg') used stand-alone (inference): [B,3,T,H,W] -> [B,1024,T/8,H/32,W/32], eval-mode BN."""
    from backbone.select_backbone import select_backbone
    from oracle import coclr_oracle as O
    net, param = select_backbone("s3dg")
    assert param["feature_size"] == 1024
    sd = {k[len("encoder_q.0."):]: v for k, v in O.with_aliases(O.synth_state(O.infonce_shapes(128, 128, network="s3dg"),
                                                                               seed=3)).items()
          if k.startswith("encoder_q.0.")}
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 8, 64, 64, generator=g).cuda()
    with torch.no_grad():
        y = net(x)
    sdd = {k: (v.cuda().double() if v.is_floating_point() else v.cuda()) for k, v in sd.items()}
    ref = O.s3d(sdd, "", x.double(), False)
    assert y.shape == (2, 1024, 1, 2, 2)
    e = _rel_l2(y, ref)
    diag["s3dg/bare_backbone_eval"] = e
    assert e < 1e-3
