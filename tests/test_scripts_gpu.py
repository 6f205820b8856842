"""The reference's command lines (main_nce.py / main_coclr.py flags, checkpoint save + --resume) drive the engine end to
end on synthetic clips: a few real training steps per case on the GPU, through the same argparse surface."""
import glob
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SMALL = ["--synthetic", "--batch_size", "4", "--seq_len", "8", "--img_dim", "64", "--moco-k", "64", "--steps-per-epoch", "3",
         "--print_freq", "1"]


def _run(script, args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, script)] + args, cwd=cwd, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("net,model", [("s3d", "infonce"), ("r50", "infonce"), ("s3d", "ubernce")])
def test_main_nce_trains_saves_resumes(tmp_path, net, model):
    # two epochs for the headline configuration (exercises checkpoint pruning), one for the others (a checkpoint of
    # the r50 model + optimizer is ~0.5 GB: keep the disk traffic of the suite small)
    E = 2 if (net, model) == ("s3d", "infonce") else 1
    out = _run("main_nce.py", SMALL + ["--net", net, "--model", model, "--epochs", str(E)], str(tmp_path))
    assert "Training from ep 0 to ep %d finished" % E in out and "loss" in out
    ck = sorted(glob.glob(str(tmp_path / "log-pretrain" / "*" / "model" / "epoch*.pth.tar")))
    assert [os.path.basename(c) for c in ck] == ["epoch%d.pth.tar" % (E - 1)]      # earlier epochs pruned (save_freq gap)
    sd = torch.load(ck[0], map_location="cpu")
    assert sd["epoch"] == E - 1 and "encoder_q.2.weight" in sd["state_dict"] and "queue" in sd["state_dict"]
    assert int(sd["state_dict"]["queue_ptr"]) == (E * 3 * 4) % 64
    out2 = _run("main_nce.py", SMALL + ["--net", net, "--model", model, "--epochs", str(E + 1), "--resume", ck[0]],
                str(tmp_path))
    assert "Training from ep %d to ep %d finished" % (E, E + 1) in out2


def test_main_nce_raw_loader_input(tmp_path):
    """--raw-input: the loader layout goes in unchanged, normalisation / clip split run inside the packing kernel."""
    out = _run("main_nce.py", SMALL + ["--net", "s3d", "--model", "infonce", "--epochs", "1", "--raw-input"], str(tmp_path))
    assert "Training from ep 0 to ep 1 finished" in out and "loss" in out and "nan" not in out.lower()


def test_main_coclr_cycle(tmp_path):
    args = ["--synthetic", "--batch_size", "4", "--seq_len", "8", "--img_dim", "64", "--moco-k", "16", "--steps-per-epoch", "7",
            "--print_freq", "1", "--net", "s3d", "--epochs", "1", "--topk", "5"]
    out = _run("main_coclr.py", args, str(tmp_path))
    assert "finished" in out and "loss" in out
    assert "queue_full False" in out and "queue_full True" in out      # the queue (16 slots) fills after 4 steps of 4 keys
