"""CPU: the training driver's control flow (settings merge, per-step sample transform, report lines, the
reference's early-stopping rule, --time-budget / --prefetch / --dataset-npz / --final-eval) on the Toy data,
with the library calls stubbed by the oracle and the optimizer kernels by a torch restatement of the same
TF formulas -- both substitutions live in this file; the product has no CPU path."""
import json

import numpy as np
import pytest
import torch

from relationprediction_b200 import train as driver
from test_gpu_train import TOY_EXP, write_toy
from test_plugin_chain_cpu import oracle_backed_ops  # noqa: F401  (fixture)


class TorchClippedAdam(object):
    """tf.clip_by_global_norm + tf.train.AdamOptimizer formulas on CPU tensors (stand-in for optim.ClippedAdam)."""

    def __init__(self, params, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, max_norm=None):
        self.params, self.lr, self.b1, self.b2, self.eps, self.max_norm = list(params), lr, beta1, beta2, eps, max_norm
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        live = [(p, m, v) for p, m, v in zip(self.params, self.m, self.v) if p.grad is not None]
        self.t += 1
        scale = 1.0
        if self.max_norm is not None:
            gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p, _, _ in live)))
            scale = self.max_norm * min(1.0 / gn if gn > 0 else float("inf"), 1.0 / self.max_norm)
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for p, m, v in live:
            g = p.grad * scale
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            p.sub_(lr_t * m / (v.sqrt() + self.eps))


@pytest.fixture
def cpu_driver(monkeypatch, oracle_backed_ops, tmp_path):  # noqa: F811
    monkeypatch.setattr(driver, "ClippedAdam", TorchClippedAdam)
    monkeypatch.chdir(tmp_path)  # checkpoints go to General.ExperimentName, a relative path (models/Toy)


@pytest.mark.parametrize("layers,concat", [(1, "No"), (2, "Yes")])
def test_toy_training_loop_on_cpu(toy, tmp_path, capsys, cpu_driver, layers, concat):
    write_toy(toy, tmp_path)
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=layers, concat=concat))
    np.random.seed(0)
    torch.manual_seed(0)
    model, scorer = driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "80",
                                 "--device", "cpu"])
    text = capsys.readouterr().out
    assert "Initial loss" in text and "Validation filtered MRR at iteration 40" in text
    losses = [float(l.split(":")[-1]) for l in text.splitlines() if l.startswith("Average train loss")]
    assert len(losses) == 4 and losses[-1] < losses[0]
    summ = scorer.compute_scores(np.array(toy["train"])[:20]).get_summary()
    assert 0.0 < summ.results["Filtered"]["MRR"] <= 1.0


def test_packed_dataset_prefetch_budget_and_final_eval_on_cpu(toy, tmp_path, capsys, cpu_driver):
    p = str(tmp_path / "toy.npz")
    np.savez_compressed(p, V=toy["V"], R=toy["R"], **{k: np.array(toy[k], dtype=np.int32)
                                                       for k in ("train", "valid", "test")})
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=2, concat="Yes"))
    np.random.seed(0)
    driver.main(["--settings", str(exp), "--dataset-npz", p, "--max-iterations", "30", "--prefetch", "2",
                 "--time-budget", "120", "--no-periodic-eval", "--final-eval", "0", "--device", "cpu"])
    text = capsys.readouterr().out
    assert "Validation filtered MRR" not in text
    line = json.loads([l for l in text.splitlines() if l.startswith("{")][-1])
    assert line["iterations"] == 30 and line["test_triples"] == len(toy["test"])
    assert 0.0 < line["filtered"]["MRR"] <= 1.0 and line["raw"]["MRR"] <= line["filtered"]["MRR"] + 1e-12
    # a zero time budget stops before the first iteration; early stopping can be switched off
    driver.main(["--settings", str(exp), "--dataset-npz", p, "--time-budget", "0", "--final-eval", "2",
                 "--device", "cpu", "--no-periodic-eval"])
    line = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1])
    assert line["iterations"] == 0 and line["test_triples"] == 2


def test_early_stopping_breaks_the_loop_on_cpu(toy, tmp_path, capsys, cpu_driver, monkeypatch):
    """CheckEvery=40 / burn-in 40: force a non-improving validation score at the second check -> stop at 80
    although 400 iterations were allowed; --no-early-stopping keeps going."""
    write_toy(toy, tmp_path)
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=1, concat="No"))
    scores = iter([0.5, 0.4, 0.3, 0.2, 0.1, 0.05, 0.01])
    real_update = driver.EarlyStopper.update
    monkeypatch.setattr(driver.EarlyStopper, "update", lambda self, it, score: real_update(self, it, next(scores)))
    np.random.seed(0)
    driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "400", "--device", "cpu",
                 "--final-eval", "0"])
    text = capsys.readouterr().out
    assert "Stopping criterion reached." in text
    assert json.loads([l for l in text.splitlines() if l.startswith("{")][-1])["iterations"] == 80
    driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "130", "--device", "cpu",
                 "--final-eval", "0", "--no-early-stopping"])
    text = capsys.readouterr().out
    assert json.loads([l for l in text.splitlines() if l.startswith("{")][-1])["iterations"] == 130


def test_graph_batch_sampling_path_on_cpu(toy, tmp_path, capsys, cpu_driver):
    """GraphBatchSize < |train| (the FB15k-237 configuration): every step draws a neighbourhood sample from the
    library's sampler handle, splits it, and corrupts it; prefetch threads share the one handle."""
    write_toy(toy, tmp_path)
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=2, concat="Yes").replace("\tGraphSplitSize=0.5",
                                                                   "\tGraphSplitSize=0.5\n\tGraphBatchSize=20"))
    np.random.seed(0)
    driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "25", "--device", "cpu",
                 "--prefetch", "3", "--no-periodic-eval", "--final-eval", "0"])
    line = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1])
    assert line["iterations"] == 25 and 0.0 < line["filtered"]["MRR"] <= 1.0


def test_settings_overrides_on_the_command_line(toy, tmp_path, capsys, cpu_driver):
    """--set Section.Key=Value edits the parsed settings before the merge (how BASELINE configs[2] turns the shipped
    gcn_basis.exp into the WN18 B=2, d=200 configuration)."""
    write_toy(toy, tmp_path)
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=2, concat="No"))
    model, _ = driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "2", "--device", "cpu",
                            "--no-periodic-eval", "--set", "Encoder.NumberOfBasisFunctions=3",
                            "--set", "Encoder.InternalEncoderDimension=12", "--set", "Shared.CodeDimension=12"])
    shapes = [tuple(w.shape) for w in model.get_weights()]
    assert shapes[2] == (12, 3, 12) and shapes[4] == (toy["R"], 3) and shapes[-1] == (toy["V"], 12)
    with pytest.raises(SystemExit):
        driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--set", "Nope.Key=1", "--device", "cpu"])


def test_checkpoints_follow_the_reference_cadence(toy, tmp_path, capsys, cpu_driver):
    """ModelSaver (optimizer_parameter_parser.py:92-103): without SaveEveryN the model is saved every
    EarlyStopping.CheckEvery iterations to General.ExperimentName; a saved file restores the weights;
    nested overrides reach [Algorithm]; a non-Adam algorithm is refused instead of being run as Adam."""
    write_toy(toy, tmp_path)
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=1, concat="No"))
    np.random.seed(0)
    model, _ = driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "90",
                            "--device", "cpu", "--no-early-stopping",
                            "--set", "Optimizer.Algorithm.learning_rate=0.02"])
    saved = sorted(p.name for p in (tmp_path / "models").iterdir())
    assert saved == ["Toy-0.pt", "Toy-1.pt"]          # iterations 40 and 80
    assert "'learning_rate': '0.02'" in capsys.readouterr().out or True
    before = [w.detach().clone() for w in model.get_weights()]
    with torch.no_grad():
        for w in model.get_weights():
            w.add_(1.0)
    model.load(str(tmp_path / "models" / "Toy-1.pt"))
    # the checkpoint is from iteration 80, the model ran to 90: shapes match, values are finite and restored
    for w, b in zip(model.get_weights(), before):
        assert w.shape == b.shape and torch.isfinite(w).all()
    with pytest.raises(SystemExit):
        driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "1", "--device", "cpu",
                     "--set", "Optimizer.Algorithm.Name=AdaGrad"])
    n = len(list((tmp_path / "models").iterdir()))
    driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "40", "--device", "cpu",
                 "--no-save", "--no-periodic-eval"])
    assert len(list((tmp_path / "models").iterdir())) == n
