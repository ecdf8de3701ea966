"""CPU (-m "not gpu"): the oracle restatement against golden vectors generated from the REAL
reference modules (oracle/make_golden.py; fixtures in tests/golden/).  fp32 CPU on both sides, so
the bar is ~1e-6 relative (bit-exact was observed for the wo-extra path)."""
import os

import pytest
import torch

from common import build_product, load_golden, oracle_forward, probe_values, relerr
from oracle import duoformer_oracle as orc
from oracle import synth

CASES = ["wo2_d12", "wo4_d2", "wo4_d12", "wo3_d2", "wo2_channel_d2", "wo2_swav_d2", "mm2_d12", "mm2_d1", "mm2_d2_b1"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(os.cpu_count() or 1)
    gold = load_golden(name)
    case = gold["case"]
    model = build_product(case)
    template = model.state_dict()
    assert sorted(template.keys()) == gold["keys"], "state_dict schema differs from the reference's"
    sd = synth.synth_state_dict(template, seed=gold["weight_seed"])
    x = synth.synth_images(case["batch"], seed=gold["input_seed"])
    cap = {}
    with torch.no_grad():
        logits = oracle_forward(case, x, sd, capture=cap)
    assert tuple(logits.shape) == tuple(gold["logits"].shape)
    assert relerr(logits, gold["logits"]) < 2e-5
    # reference hands tokens WITHOUT pos_embed_for_scale to the transformer
    tok_in = cap["tokens"] - sd["vision_transformer.pos_embed_for_scale"]
    pr = gold["probes"]["tokens_in"]
    assert tuple(tok_in.shape) == pr["shape"]
    assert (probe_values(tok_in, pr) - pr["values"]).abs().max().item() <= 2e-5 * pr["absmax"]
    for key, pr in gold["probes"].items():
        if key == "tokens_in":
            continue
        t = cap[key]
        assert tuple(t.shape) == pr["shape"], key
        assert (probe_values(t, pr) - pr["values"]).abs().max().item() <= 2e-5 * pr["absmax"], key
        assert abs(t.float().norm().item() - pr["norm"]) <= 2e-5 * pr["norm"], key


def test_logits_depend_on_input_and_weights():
    """The synthetic weight set must not be degenerate: different images give
    different logits, by much more than fp32 noise."""
    gold = load_golden("wo2_d12")
    lg = gold["logits"]
    assert (lg[0] - lg[1]).abs().max().item() > 1e-3
    assert lg.abs().max().item() > 0.5


def test_index_table_matches_reference_literals_g7():
    # spot values read off the reference's literal tables (model_wo_extra_params.py:117-212)
    t2 = orc.index_table(2, 7)
    assert t2[0].tolist() == [0, 14, 1, 15]
    assert t2[8].tolist() == [2 * 14 + 2, 3 * 14 + 2, 2 * 14 + 3, 3 * 14 + 3]
    t1 = orc.index_table(1, 7)
    assert t1[0].tolist() == [0, 1, 2, 3, 28, 29, 30, 31, 56, 57, 58, 59, 84, 85, 86, 87]
    t0 = orc.index_table(0, 7)
    assert t0[48, 0].item() == 8 * 6 * 56 + 8 * 6 and t0[48, 63].item() == (8 * 6 + 7) * 56 + 8 * 6 + 7
    assert orc.index_table(3, 7)[:, 0].tolist() == list(range(49))
