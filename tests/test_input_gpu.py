"""Fused input staging: the reference's GPU-side `tr` -- T.Normalize(mean, std, channel=1), view to
[B,3,num_seq,seq_len,H,W], transpose(1,2), contiguous (main_nce.py:207-209,299-302; utils/transforms.py:57-63) -- folded
into the clip-packing kernel.  model(raw loader tensor) must equal model(tr(raw)) with tr done by torch ops."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
pytestmark = pytest.mark.gpu

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _tr(x, num_seq, seq_len):
    """The reference's transform, restated with the same torch ops (normalize of utils/transforms.py:57-63, channel=1)."""
    shape = [1] * x.dim()
    shape[1] = -1
    mean = torch.as_tensor(MEAN).to(x.device).reshape(shape)
    std = torch.as_tensor(STD).to(x.device).reshape(shape)
    y = (x - mean) / std
    B, _, _, H, W = x.shape
    return y.view(B, 3, num_seq, seq_len, H, W).transpose(1, 2).contiguous()


def _model(K=64):
    from model.pretrain import InfoNCE
    from oracle import coclr_oracle as O
    torch.manual_seed(0)
    m = InfoNCE("s3d", 128, K, 0.999, 0.07)
    m.load_state_dict(O.with_aliases(O.synth_state(O.infonce_shapes(128, K), seed=0, ptr=0)), strict=True)
    return m.cuda().train()


@pytest.mark.parametrize("hw", [64, 56])
def test_fused_transform_equals_reference_tr(hw, diag):
    from coclr_b200 import moco
    B, T = 4, 8
    g = torch.Generator().manual_seed(11)
    raw = torch.rand(B, 3, 2 * T, hw, hw, generator=g).cuda()
    a, b = _model(), _model()
    a.set_input_transform(MEAN, STD)
    torch.manual_seed(5)
    la, _ = a(raw)                          # loader layout in, transform fused
    torch.manual_seed(5)
    lb, _ = b(_tr(raw, 2, T))               # reference-style pre-transformed block
    torch.cuda.synchronize()
    pa = a.encoder_q._engine.plans[(B, T, hw, hw, True, True)].input.pl
    pb = b.encoder_q._engine.plans[(B, T, hw, hw, True, True)].input.pl
    assert torch.equal(pa.hi, pb.hi) and torch.equal(pa.lo, pb.lo)      # same bits reach the first conv
    ka = a.encoder_k._engine.plans[(B, T, hw, hw, True, False)].input.pl
    kb = b.encoder_k._engine.plans[(B, T, hw, hw, True, False)].input.pl
    assert torch.equal(ka.hi, kb.hi) and torch.equal(ka.lo, kb.lo)      # ... and the shuffled key clips
    err = float((la - lb).abs().max() / lb.abs().max())
    diag["input/fused_tr_logits_rel_%d" % hw] = err
    assert err < 1e-5
    # gradients flow through the fused path as well
    moco.nce_cross_entropy(la, torch.zeros(B, dtype=torch.long, device="cuda")).backward()
    assert float(a.encoder_q[0].Conv_1a.conv1.weight.grad.abs().sum()) > 0


def test_six_dim_block_is_normalised_too():
    B, T, hw = 2, 8, 64
    g = torch.Generator().manual_seed(12)
    raw = torch.rand(B, 3, 2 * T, hw, hw, generator=g).cuda()
    blk = raw.view(B, 3, 2, T, hw, hw).transpose(1, 2).contiguous()      # un-normalised [B,2,3,T,H,W]
    a, b = _model(), _model()
    a.set_input_transform(MEAN, STD)
    torch.manual_seed(5)
    la, _ = a(blk)
    torch.manual_seed(5)
    lb, _ = b(_tr(raw, 2, T))
    assert float((la - lb).abs().max() / lb.abs().max()) < 1e-5
