"""World-size-2 `gloo` test (CPU) of the data-parallel host logic of the MoCo step: rank-0 permutation
broadcast, shuffle-BN gather, the single merged key all-gather (un-shuffle + enqueue), queue pointer mirror.
The CUDA kernels are replaced by torch stand-ins (the encoder by a deterministic per-clip function), so what is
checked is exactly the cross-rank plumbing, against the oracle's simulated-world restatement of
model/pretrain.py:82-143."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_encode(x):
    """Deterministic stand-in for encoder_k: a normalised 128-vector that depends only on the clip content."""
    feat = x.flatten(1)[:, :128 * 3].view(x.shape[0], 128, 3).sum(-1) + x.flatten(1).mean(1, keepdim=True)
    return torch.nn.functional.normalize(feat, dim=1)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from model.pretrain import InfoNCE
    from coclr_b200 import moco
    torch.manual_seed(0)
    B, K = 4, 32
    m = InfoNCE("s3d", 128, K)
    m.encoder_k.encode = lambda x, batch_index=None, batch=None, peers=None: _fake_encode(x[batch_index] if batch_index is not None else x)
    moco.enqueue = lambda queue, keys, ptr: queue.__setitem__((slice(None), slice(ptr, ptr + keys.shape[0])), keys.T)
    g = torch.Generator().manual_seed(100 + rank)
    x2 = torch.randn(B, 3, 2, 8, 8, generator=g)
    torch.manual_seed(5 + rank)          # ranks draw DIFFERENT permutations; only rank 0's may be used
    k_local, k_global = m._shuffled_keys(x2)
    ptr0 = m._dequeue_and_enqueue(k_global)
    ptr1 = m._ptr()
    k2_local, k2_global = m._shuffled_keys(x2 * 2)
    m._dequeue_and_enqueue(k2_global)
    torch.save({"x2": x2, "k_local": k_local, "k_global": k_global, "queue": m.queue.clone(), "ptr0": ptr0, "ptr1": ptr1,
                "ptr2": m._ptr(), "queue_ptr": m.queue_ptr.clone()}, os.path.join(out, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_shuffle_gather_enqueue_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % i)) for i in range(world)]
    B = 4
    # every rank ends with identical global keys / queue / pointer (replica consistency)
    assert torch.equal(r[0]["k_global"], r[1]["k_global"])
    assert torch.equal(r[0]["queue"], r[1]["queue"])
    assert r[0]["ptr0"] == r[1]["ptr0"] == 0 and r[0]["ptr1"] == r[1]["ptr1"] == world * B
    assert r[0]["ptr2"] == 2 * world * B and int(r[0]["queue_ptr"]) == 2 * world * B
    # un-shuffled keys are in ORIGINAL rank-major order: key i of rank r == encode(x2 of rank r)[i]
    for i in range(world):
        want = _fake_encode(r[i]["x2"])
        assert torch.allclose(r[i]["k_local"], want, atol=1e-6)
        assert torch.allclose(r[0]["k_global"][i * B:(i + 1) * B], want, atol=1e-6)
    # queue columns [0, 8) hold the first step's global keys, [8, 16) the second step's
    assert torch.allclose(r[0]["queue"][:, :world * B], r[0]["k_global"].T, atol=1e-6)
    second = torch.cat([_fake_encode(r[i]["x2"] * 2) for i in range(world)])
    assert torch.allclose(r[0]["queue"][:, world * B:2 * world * B], second.T, atol=1e-6)


def test_shuffle_matches_oracle_simulated_world():
    """Same index algebra as the oracle's simulated world (which is pinned to the reference in test_oracle.py)."""
    torch.manual_seed(3)
    W, B = 4, 3
    idx_shuffle = torch.randperm(W * B)
    idx_unshuffle = torch.argsort(idx_shuffle)
    x_all = torch.randn(W * B, 3, 2, 8, 8)
    k_sh = [_fake_encode(x_all[idx_shuffle.view(W, -1)[r]]) for r in range(W)]
    k_global = torch.cat(k_sh, 0)[idx_unshuffle]
    assert torch.allclose(k_global, _fake_encode(x_all), atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# UberNCE / CoCLR at world size 2: label / video-id / second-view key gathers and the integer queues
# (model/pretrain.py:211-227, 321-341) -- host logic only, kernels replaced by torch stand-ins as above
# ---------------------------------------------------------------------------------------------------------------
def _stub_kernels(m, moco):
    for enc in (m.encoder_q, m.encoder_k, getattr(m, "sampler", None)):
        if enc is not None:
            enc.encode = lambda x, batch_index=None, batch=None, peers=None: _fake_encode(x[batch_index] if batch_index is not None else x)
    moco.enqueue = lambda queue, keys, ptr: queue.__setitem__((slice(None), slice(ptr, ptr + keys.shape[0])), keys.T)
    moco.momentum_update = lambda q, k, mm: None
    moco.nce_logits = lambda q, k, queue, T: torch.cat([(q * k).sum(1, keepdim=True), q @ queue.clone()], 1) / T

    def mask_topk(kf, queue_second, k_vsource, queue_vname, topk):     # the reference's statements (pretrain.py:392-413)
        src = k_vsource.unsqueeze(1) == queue_vname.unsqueeze(0)
        mask = src.clone()
        if topk:
            sim = kf.matmul(queue_second.clone())
            sim[src] = -float("inf")
            _, idx = torch.topk(sim, topk, dim=1)
            onehot = torch.zeros_like(sim)
            onehot.scatter_(1, idx, 1)
            mask[onehot.bool()] = True
        return torch.cat([torch.ones((mask.shape[0], 1), dtype=torch.bool), mask], dim=1)
    moco.mask_topk = mask_topk


def _worker_ext(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from model import pretrain as P
    from coclr_b200 import moco
    B, K = 4, 32
    g = torch.Generator().manual_seed(300 + rank)
    res = {}
    # ---- UberNCE: forward() with the CUDA guard lifted (the stand-ins run on the CPU) ----
    torch.manual_seed(0)
    u = P.UberNCE("s3d", 128, K)
    _stub_kernels(u, moco)
    u.queue_label[:16] = torch.arange(16) % 5
    block = torch.randn(B, 2, 3, 2, 8, 8, generator=g)
    label = torch.randint(0, 5, (B,), generator=g)
    def qk_cpu(self, block):       # InfoNCE._qk without its CUDA stream handling (the stand-ins run on the CPU)
        x1, x2 = self._views(block)
        with torch.no_grad():
            self._momentum_update_key_encoder()
            k, k_global = self._shuffled_keys(x2)
        return self.encoder_q.encode(x1), k, k_global, True
    P.InfoNCE._qk = qk_cpu
    if True:
        torch.manual_seed(11 + rank)
        logits, mask = u(block, label)
        res["u_mask"], res["u_queue_label"], res["u_ptr"] = mask.clone(), u.queue_label.clone(), int(u.queue_ptr)
        res["u_label"], res["u_logits"] = label, logits.detach().clone()
        # ---- CoCLR ----
        torch.manual_seed(0)
        c = P.CoCLR("s3d", 128, K, topk=3)
        _stub_kernels(c, moco)
        c.encoder_q.encode = lambda x, **kw: _fake_encode(x) * torch.ones(1, requires_grad=True)   # q must require grad (train mode)

        def qkf_cpu(x1, x2, f2):      # CoCLR._qkf without its CUDA stream handling
            with torch.no_grad():
                c._momentum_update_key_encoder()
                k, k_global = c._shuffled_keys(x2)
                kf = c.sampler.encode(f2)
            return c.encoder_q.encode(x1), k, k_global, kf, True
        c._qkf = qkf_cpu
        c.queue_label[:] = 1                                     # queue full -> top-k branch
        c.queue_vname[:] = torch.arange(K) % 7
        b1 = torch.randn(B, 2, 3, 2, 8, 8, generator=g)
        b2 = torch.randn(B, 2, 3, 2, 8, 8, generator=g)
        vname = torch.randint(0, 7, (B,), generator=g)
        torch.manual_seed(21 + rank)
        lg, mk = c(b1, b2, vname)
        res.update(c_mask=mk.clone(), c_queue_vname=c.queue_vname.clone(), c_queue_second=c.queue_second.clone(),
                   c_queue=c.queue.clone(), c_ptr=int(c.queue_ptr), c_vname=vname, c_f2=b2[:, 1].clone(), c_x2=b2[:, 0].clone())
    torch.save(res, os.path.join(out, "ext%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_ubernce_coclr_queues_world2(tmp_path):
    world, B, K = 2, 4, 32
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker_ext, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(str(tmp_path), "ext%d.pt" % i)) for i in range(world)]
    # UberNCE (pretrain.py:211-227): labels of ALL ranks enter the label queue in rank-major order on every rank
    assert torch.equal(r[0]["u_queue_label"], r[1]["u_queue_label"]) and r[0]["u_ptr"] == r[1]["u_ptr"] == world * B
    assert torch.equal(r[0]["u_queue_label"][:world * B], torch.cat([r[0]["u_label"], r[1]["u_label"]]))
    old = torch.full((K,), -1, dtype=torch.long)
    old[:16] = torch.arange(16) % 5
    for i in range(world):        # the mask uses the label queue as it was BEFORE this step's enqueue (pretrain.py:271)
        want = torch.cat([torch.ones(B, 1, dtype=torch.bool), r[i]["u_label"][:, None] == old[None, :]], 1)
        assert torch.equal(r[i]["u_mask"], want)
    # CoCLR (pretrain.py:321-341): video ids and second-view keys of all ranks, rank-major, identical replicas
    for key in ("c_queue_vname", "c_queue_second", "c_queue"):
        assert torch.equal(r[0][key], r[1][key]), key
    assert r[0]["c_ptr"] == r[1]["c_ptr"] == world * B
    assert torch.equal(r[0]["c_queue_vname"][:world * B], torch.cat([r[0]["c_vname"], r[1]["c_vname"]]))
    kf = torch.cat([_fake_encode(r[i]["c_f2"]) for i in range(world)])
    assert torch.allclose(r[0]["c_queue_second"][:, :world * B], kf.T, atol=1e-6)
    k = torch.cat([_fake_encode(r[i]["c_x2"]) for i in range(world)])
    assert torch.allclose(r[0]["c_queue"][:, :world * B], k.T, atol=1e-6)
    for i in range(world):        # same-source positives + topk=3 mined positives + the self column
        m = r[i]["c_mask"]
        src = r[i]["c_vname"][:, None] == (torch.arange(K) % 7)[None, :]
        assert bool(m[:, 0].all()) and bool((m[:, 1:] | ~src).all())
        assert torch.equal(m[:, 1:].sum(1), src.sum(1) + 3)
