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
