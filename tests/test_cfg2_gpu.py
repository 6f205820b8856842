"""GPU parity at the shapes bench.py times (BASELINE.json config 2: S3D InfoNCE, bs 32, 32 x 128^2 clips, K=2048) and
the fused NCE kernels at config 3's queue length (K=16384).

The small-shape tests (test_infonce_gpu.py, config 1) run one or two tiles per layer; here every layer runs its
production tile counts, weight-gradient split counts, statistics flush mode and pipeline depths.  Truth: the oracle
(oracle/coclr_oracle.py, pinned bit-for-bit to the unmodified reference by tests/test_oracle.py) evaluated in float64
on the same device from the same state and inputs; tolerance 1e-3 relative on logits / loss / queue (north star), the
oracle-relative budget of SURVEY.md section 7 on gradients (the reference's own fp32 arithmetic is ~2e-2 off the
float64 gradient of this saturated loss)."""
import numpy as np
import pytest
import torch

from test_infonce_gpu import _build, _oracle_truth, _rel, _rel_l2

pytestmark = pytest.mark.gpu

B, T, HW, K = 32, 32, 128, 2048


@pytest.fixture(scope="module")
def step2():
    from coclr_b200 import moco
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model, sd = _build(K=K, ptr=K - B)          # the enqueue fills the last B columns and wraps the pointer to 0
    g = torch.Generator(device="cuda").manual_seed(2024)
    block = torch.randn(B, 2, 3, T, HW, HW, device="cuda", generator=g)
    torch.manual_seed(99)
    idx = torch.randperm(B)
    torch.manual_seed(99)                        # the module draws the same permutation from the CPU RNG
    logits, labels = model(block)
    loss = moco.nce_cross_entropy(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    rec64 = {}
    truth = _oracle_truth(sd, block, idx, torch.float64, rec64)
    # the float64 records of the 77 activations at B=32 are ~7 GB: fine on a 180 GB part
    return dict(model=model, sd=sd, logits=logits, labels=labels, loss=loss, truth=truth, rec64=rec64, block=block,
                idx=idx)


def test_cfg2_forward_loss_queue(step2, diag):
    model, logits, loss = step2["model"], step2["logits"], step2["loss"]
    sd64, lg64, loss64 = step2["truth"]
    e_logits = _rel(logits, lg64)
    e_loss = abs(float(loss) - float(loss64)) / max(1.0, abs(float(loss64)))
    e_queue = _rel(model.queue, sd64["queue"])
    diag["cfg2/logits_vs_fp64"] = e_logits
    diag["cfg2/loss"] = [float(loss), float(loss64)]
    diag["cfg2/queue_vs_fp64"] = e_queue
    assert logits.shape == (B, 1 + K)
    assert e_logits < 1e-3 and e_loss < 1e-3 and e_queue < 1e-3
    assert int(model.queue_ptr) == int(sd64["queue_ptr"]) == 0
    # only the last B columns changed
    q0 = step2["sd"]["queue"].cuda()
    assert torch.equal(model.queue[:, :K - B], q0[:, :K - B])
    # EMA is exact (same two-rounding formula), BN running statistics follow the batch statistics
    k = "encoder_k.0.Conv_2c.conv1.weight"
    assert _rel(model.state_dict()[k], sd64[k]) < 1e-6
    for k in ("encoder_q.0.Conv_1a.bn1.running_var", "encoder_k.0.Mixed_4d.branch1.1.bn2.running_mean"):
        assert _rel(model.state_dict()[k], sd64[k]) < 1e-3, k


def test_cfg2_layerwise(step2, diag):
    """All 77 post-BN-ReLU activations of encoder_q at the benchmarked shape against the float64 oracle."""
    model, rec = step2["model"], step2["rec64"]
    eng = model.encoder_q._engine
    plan = eng.plans[(B, T, HW, HW, True, True)]
    per, worst = {}, 0.0
    for t in eng.graph.tensors:
        if not t.pending:
            continue
        y = plan.acts[t.index].pl.value()
        for nm, coff, c in t.bn_members:
            ref = rec["encoder_q." + nm].permute(0, 2, 3, 4, 1)
            e = _rel_l2(y[..., coff:coff + c].double(), ref)
            per[nm] = e
            worst = max(worst, e)
    diag["cfg2/layerwise_max_rel_l2"] = worst
    diag["cfg2/layerwise"] = per
    assert len(per) == 77
    assert worst < 1e-3, sorted(per.items(), key=lambda kv: -kv[1])[:5]


def test_cfg2_gradients(step2, diag):
    """Gradient budget against the float64 oracle, with the fp32 oracle's own deviation as the yardstick."""
    model = step2["model"]
    sd64 = step2["truth"][0]
    sd32 = _oracle_truth(step2["sd"], step2["block"], step2["idx"], torch.float32)[0]
    named = dict(model.named_parameters())
    out, bad = {}, []
    for k in sorted(sd64):
        if not k.startswith("encoder_q.") or not (k.endswith(".weight") or k.endswith(".bias")):
            continue
        g64 = sd64[k].grad
        e_new, e_ref = _rel_l2(named[k].grad, g64), _rel_l2(sd32[k].grad, g64)
        out[k] = [e_new, e_ref]
        if not e_new < max(5 * e_ref, 3e-2):
            bad.append((k, e_new, e_ref))
    med_new = float(np.median([v[0] for v in out.values()]))
    med_ref = float(np.median([v[1] for v in out.values()]))
    diag["cfg2/grad_median_new_ref"] = [med_new, med_ref]
    diag["cfg2/grad_err_new_vs_ref"] = out
    diag["cfg2/grad_exceptions"] = bad
    assert len(out) == 235
    assert med_new < 1.5 * med_ref + 1e-3, (med_new, med_ref)      # measured 2.6e-2 vs 2.0e-2
    assert len(bad) <= 2, bad[:8]


@pytest.mark.parametrize("Kq,Bq", [(16384, 32), (16384, 256), (2048, 32)])
def test_nce_kernels_long_queue(Kq, Bq, diag):
    """coclr_nce_logits_ce / coclr_nce_logits_bwd at config 3's queue length (K=16384; bs 32 per GPU) against
    torch float64: logits, per-row loss, d(mean loss)/d(logits) and dq (model/pretrain.py:175-182, main_nce.py:201,314)."""
    import ctypes as C
    from coclr_b200 import lib as L
    g = torch.Generator(device="cuda").manual_seed(5)
    D, Tt = 128, 0.07
    q = torch.nn.functional.normalize(torch.randn(Bq, D, device="cuda", generator=g), dim=1)
    k = torch.nn.functional.normalize(q + 0.3 * torch.randn(Bq, D, device="cuda", generator=g), dim=1)
    queue = torch.nn.functional.normalize(torch.randn(D, Kq, device="cuda", generator=g), dim=0)
    logits = torch.empty(Bq, Kq + 1, device="cuda")
    loss_rows = torch.empty(Bq, device="cuda")
    dlogits = torch.empty(Bq, Kq + 1, device="cuda")
    dq = torch.empty(Bq, D, device="cuda")
    lib = L.load()
    ws = torch.empty(2 * Bq * ((Kq + 1023) // 1024), device="cuda")       # long queues: K slices on separate CTAs
    L.check(lib.coclr_nce_logits_ce(L.dptr(q), L.dptr(k), L.dptr(queue), Tt, Bq, D, Kq, L.dptr(logits), L.dptr(loss_rows),
                                    L.dptr(dlogits), L.dptr(ws), L.stream_ptr()), "coclr_nce_logits_ce")
    L.check(lib.coclr_nce_logits_bwd(L.dptr(dlogits), L.dptr(k), L.dptr(queue), Tt, Bq, D, Kq, L.dptr(dq),
                                     L.stream_ptr()), "coclr_nce_logits_bwd")
    torch.cuda.synchronize()
    qd = q.double().requires_grad_(True)
    lg = torch.cat([(qd * k.double()).sum(1, keepdim=True), qd @ queue.double()], 1) / Tt
    lg.retain_grad()
    rows = torch.logsumexp(lg, 1) - lg[:, 0]
    rows.mean().backward()
    e = {"logits": _rel(logits, lg), "loss_rows": _rel(loss_rows, rows), "dlogits": _rel(dlogits, lg.grad),
         "dq": _rel(dq, qd.grad)}
    diag["nce/K%d_B%d" % (Kq, Bq)] = e
    assert e["logits"] < 1e-5 and e["loss_rows"] < 1e-5 and e["dlogits"] < 1e-4 and e["dq"] < 1e-4, e
