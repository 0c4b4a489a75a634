"""Host-side control flow of the reference-compatible CLI (dlrm_b200/cli.py) on CPU: the model and the
fused optimizers are replaced by stand-ins, so what runs here is flag handling, the batch sources
(uniform / gaussian / trace-driven), the per-epoch re-seeding and the printed line format."""
import os
import re

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture
def cli_on_cpu(monkeypatch):
    import dlrm_b200.cli as cli
    import dlrm_b200.dlrm_net as dn
    import dlrm_b200.optim as fo

    seen = []

    class StandIn(torch.nn.Module):
        def __init__(self, m_spa, ln_emb, ln_bot, ln_top, **kw):
            super().__init__()
            self.lin = torch.nn.Linear(int(ln_bot[0]), 1)
            self.loss_fn = torch.nn.MSELoss()
            self.n_tables = len(ln_emb)

        def forward(self, X, lS_o, lS_i):
            assert lS_o.shape == (self.n_tables, X.shape[0]) and len(lS_i) == self.n_tables
            seen.append((X.clone(), lS_o.clone(), [i.clone() for i in lS_i]))
            return torch.sigmoid(self.lin(X))

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: self)
    monkeypatch.setattr(dn, "DLRM_Net", StandIn)
    monkeypatch.setattr(fo, "SGD", torch.optim.SGD)
    monkeypatch.setattr(fo, "RWSAdagrad", torch.optim.SGD)
    return cli, seen


BASE = ["--arch-sparse-feature-size=16", "--arch-embedding-size=64-16", "--arch-mlp-bot=5-16", "--arch-mlp-top=8-1",
        "--mini-batch-size=8", "--print-freq=1", "--use-gpu"]


def test_epochs_replay_the_same_batches_and_lines_have_the_reference_format(cli_on_cpu, capsys):
    cli, seen = cli_on_cpu
    cli.run(BASE + ["--data-size=20", "--nepochs=2", "--numpy-rand-seed=5"])
    out = capsys.readouterr().out
    assert "Using 1 GPU(s)..." in out
    lines = re.findall(r"Finished training it (\d+)/3 of epoch (\d), -1.00 ms/it, loss \d+\.\d{6}", out)
    assert lines == [("1", "0"), ("2", "0"), ("3", "0"), ("1", "1"), ("2", "1"), ("3", "1")]
    assert len(seen) == 6 and seen[2][0].shape[0] == 4          # 20 samples = 8 + 8 + 4
    for a, b in zip(seen[:3], seen[3:]):                           # RandomDataset re-seeds at batch 0
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert all(torch.equal(x, y) for x, y in zip(a[2], b[2]))


def test_gaussian_and_trace_driven_sources_and_inference_only(cli_on_cpu, capsys, monkeypatch):
    cli, seen = cli_on_cpu
    cli.run(BASE + ["--num-batches=2", "--rand-data-dist=gaussian", "--rand-data-max=60", "--rand-data-sigma=9",
                    "--inference-only"])
    out = capsys.readouterr().out       # the reference's inference-only run is ONE test pass (dlrm_s_pytorch.py:1781-1792)
    assert "Testing for inference only" in out and " accuracy " in out
    assert len(seen) == 2
    assert all(int(i.max()) <= 60 for _, _, ids in seen for i in ids)
    del seen[:]
    monkeypatch.chdir(GOLD)
    cli.run(BASE + ["--num-batches=2", "--data-generation=synthetic", "--data-trace-file=datagen_dist_emb_j.log"])
    assert len(seen) == 2
    assert all(int(ids[0].max()) < 64 and int(ids[1].max()) < 16 for _, _, ids in seen)


@pytest.mark.parametrize("flags,msg", [(["--data-generation=dataset"], "--data-generation=dataset is not supported"),
                                       (["--qr-flag"], "--qr-flag is outside"),
                                       (["--optimizer=adagrad"], "--optimizer=adagrad is not supported"),
                                       (["--arch-sparse-feature-size=8"], "does not match last dim of bottom mlp")])
def test_unsupported_options_exit_with_an_error_string(cli_on_cpu, flags, msg):
    cli, _ = cli_on_cpu
    with pytest.raises(SystemExit) as e:
        cli.run(BASE + ["--num-batches=1"] + flags)
    assert msg in str(e.value)


def test_test_pass_checkpoint_and_resume_control_flow(cli_on_cpu, capsys, tmp_path, monkeypatch):
    """--test-freq / --save-model / --load-model on the stand-in model: the reference's lines in the reference's order
    (dlrm_s_pytorch.py:1640-1715), its checkpoint key set, and a resume that skips the batches already trained while
    still drawing them (the generator's order is the reference's)."""
    cli, seen = cli_on_cpu
    orig_load = torch.load
    monkeypatch.setattr(torch, "load", lambda f, map_location=None, **k: orig_load(f, map_location="cpu", **k))   # no GPU here
    ck = str(tmp_path / "m.pt")
    common = BASE + ["--num-batches=4", "--round-targets=True", "--numpy-rand-seed=3"]
    cli.run(common + ["--test-freq=2", "--save-model=" + ck])
    out = capsys.readouterr().out.splitlines()
    body = [ln for ln in out if re.match(r"Finished|Testing at| accuracy|Saving model", ln)]
    kinds = [ln.split()[0] for ln in body]
    assert kinds == ["Finished", "Finished", "Testing", "accuracy", "Saving", "Finished", "Finished", "Testing", "accuracy",
                     "Saving"]
    assert body[2] == "Testing at - 2/4 of epoch 0," and re.fullmatch(r" accuracy \d+\.\d{3} %, best \d+\.\d{3} %", body[3])
    assert len(seen) == 4 + 2 * 4                         # 4 training batches + two test passes over 4 batches
    sd = torch.load(ck, weights_only=False)
    assert set(sd) == {"epoch", "iter", "nepochs", "nbatches", "nbatches_test", "state_dict", "opt_state_dict",
                       "train_loss", "total_loss", "test_acc"}
    assert (sd["epoch"], sd["iter"], sd["nepochs"], sd["nbatches"], sd["nbatches_test"]) == (0, 4, 1, 4, 4)
    # resume: everything of epoch 0 was trained -> every batch is drawn and skipped
    del seen[:]
    cli.run(common + ["--load-model=" + ck])
    out = capsys.readouterr().out
    assert "Loading saved model " + ck in out and "Saved at: epoch = 0/1, batch = 4/4, ntbatch = 4" in out
    assert "Finished training" not in out and len(seen) == 0
    # inference only from the checkpoint: ONE test pass, no training lines
    cli.run(common + ["--load-model=" + ck, "--inference-only"])
    out = capsys.readouterr().out
    assert "Testing for inference only" in out and " accuracy " in out and "Finished" not in out
    assert len(seen) == 4
