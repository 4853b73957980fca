"""GPU: the training driver end to end on the Toy dataset (BASELINE configs[0]: 1-layer basis R-GCN +
DistMult, no GraphBatchSize -> the whole train set is the graph batch, train.py:228-230) and a 2-layer
block model: the loss goes down and the evaluation runs."""
import numpy as np
import pytest
import torch

from relationprediction_b200 import train as driver

pytestmark = pytest.mark.gpu

TOY_EXP = """[Encoder]
\tName=gcn_basis
\tDropoutKeepProbability=0.8
\tInternalEncoderDimension=16
\tNumberOfBasisFunctions=2
\tNumberOfLayers={layers}
\tUseInputTransform=Yes
\tUseOutputTransform=No
\tAddDiagonal=No
\tDiagonalCoefficients=No
\tSkipConnections=None
\tStoreEdgeData=No
\tRandomInput=No
\tPartiallyRandomInput=No
\tConcatenation={concat}

[Decoder]
\tName=bilinear-diag
\tRegularizationParameter=0.01

[Shared]
\tCodeDimension=16

[Optimizer]
\tMaxGradientNorm=1
\tReportTrainLossEvery=20

\t[EarlyStopping]
\t\tCheckEvery=40
\t\tBurninPhaseDuration=40

\t[Algorithm]
\t\tName=Adam
\t\tlearning_rate=0.01

[General]
\tNegativeSampleRate=10
\tGraphSplitSize=0.5
\tExperimentName=models/Toy

[Evaluation]
\tMetric=MRR
"""


def write_toy(toy, tmp_path):
    ent = {int(k): v for k, v in toy["entities"].items()}
    rel = {int(k): v for k, v in toy["relations"].items()}
    (tmp_path / "entities.dict").write_text("".join("%d\t%s\n" % kv for kv in sorted(ent.items())))
    (tmp_path / "relations.dict").write_text("".join("%d\t%s\n" % kv for kv in sorted(rel.items())))
    for split in ("train", "valid", "test"):
        (tmp_path / (split + ".txt")).write_text(
            "".join("%s\t%s\t%s\n" % (ent[s], rel[r], ent[o]) for s, r, o in toy[split]))


@pytest.mark.parametrize("layers,concat", [(1, "No"), (2, "Yes")])
def test_toy_training_runs_and_learns(toy, tmp_path, capsys, layers, concat):
    write_toy(toy, tmp_path)
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=layers, concat=concat))
    np.random.seed(0)
    torch.manual_seed(0)
    model, scorer = driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "80",
                                 "--save-path", str(tmp_path / "ckpt" / "Toy")])
    text = capsys.readouterr().out
    assert "Initial loss" in text and "Validation filtered MRR" in text
    losses = [float(l.split(":")[-1]) for l in text.splitlines() if l.startswith("Average train loss")]
    assert len(losses) == 4 and losses[-1] < losses[0]
    summ = scorer.compute_scores(np.array(toy["train"])[:20]).get_summary()
    assert 0.0 < summ.results["Filtered"]["MRR"] <= 1.0


def test_packed_dataset_prefetch_and_final_eval(toy, tmp_path, capsys):
    """The round-trip a real-dataset run uses: --dataset-npz, background sample threads, a time budget and the
    final JSON line with raw / filtered ranking metrics."""
    import json
    p = str(tmp_path / "toy.npz")
    np.savez_compressed(p, V=toy["V"], R=toy["R"], **{k: np.array(toy[k], dtype=np.int32)
                                                       for k in ("train", "valid", "test")})
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=2, concat="Yes"))
    np.random.seed(0)
    driver.main(["--settings", str(exp), "--dataset-npz", p, "--max-iterations", "60", "--prefetch", "2",
                 "--time-budget", "60", "--no-periodic-eval", "--final-eval", "0", "--no-save"])
    text = capsys.readouterr().out
    assert "Validation filtered MRR" not in text
    line = json.loads([l for l in text.splitlines() if l.startswith("{")][-1])
    assert line["iterations"] == 60 and line["test_triples"] == len(toy["test"])
    assert 0.0 < line["filtered"]["MRR"] <= 1.0 and line["raw"]["MRR"] <= line["filtered"]["MRR"] + 1e-12
