"""CPU: a reference checkpoint's optimizer state (per-tensor Adam moments in the reference's parameter order,
model/avatar_model.py:148-155,163-176) is re-laid out into the flat parameter buffer; the order is pinned to a fixture generated
from the reference's own module (oracle/gen_golden.py: gen_param_order)."""
import json
import os
import types

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref_order():
    return json.load(open(os.path.join(GOLD, "pop_param_order.json")))


def test_reference_param_order_matches_fixture():
    from gaussianavatar_b200.network import REFERENCE_PARAM_ORDER
    assert list(REFERENCE_PARAM_ORDER) == [n for n, _ in _ref_order()]


def test_reference_optimizer_state_translation_roundtrip():
    from gaussianavatar_b200.avatar_model import AvatarModel
    from gaussianavatar_b200.network import POP_no_unet
    net = POP_no_unet(c_geom=64, hsize=128)
    geo = torch.nn.Parameter(torch.zeros(1, 64, 16, 16))
    order = _ref_order()
    g = torch.Generator().manual_seed(0)
    # what torch.optim.Adam over the reference's [net params..., geo_feature] would have saved after 7 steps
    ref_params = [torch.nn.Parameter(torch.randn(*shape, generator=g)) for _, shape in order]
    ref_opt = torch.optim.Adam([{"params": ref_params, "lr": 3e-3}, {"params": [geo], "lr": 5e-4}])
    for _ in range(7):
        for p in ref_params + [geo]:
            p.grad = torch.randn(p.shape, generator=g)
        ref_opt.step()
    osd = ref_opt.state_dict()
    holder = types.SimpleNamespace(net=net)
    new = AvatarModel.translate_optimizer_state(holder, osd)
    assert sorted(new["state"].keys()) == [0, 1] and [gr["params"] for gr in new["param_groups"]] == [[0], [1]]
    assert float(new["state"][0]["step"]) == 7.0 and new["param_groups"][0]["lr"] == 3e-3 and new["param_groups"][1]["lr"] == 5e-4
    for key in ("exp_avg", "exp_avg_sq"):
        flat = new["state"][0][key]
        assert flat.shape == net.flat.shape
        back = net.reference_tensors_from_flat(flat)
        for (name, _), t, i in zip(order, back, range(len(order))):
            assert torch.equal(t, osd["state"][i][key]), name
        # everything that is not a reference parameter (alignment / 66->72 padding) has zero moments
        used = sum(int(torch.tensor(shape).prod()) for _, shape in order)
        assert int((flat != 0).sum()) <= used
    assert torch.equal(new["state"][1]["exp_avg"], osd["state"][len(order)]["exp_avg"])
    # our optimizer accepts it
    ours = torch.optim.Adam([{"params": [net.flat], "lr": 1.0}, {"params": [geo], "lr": 1.0}])
    ours.load_state_dict(new)
    assert ours.param_groups[0]["lr"] == 3e-3
    assert torch.equal(ours.state[net.flat]["exp_avg"], new["state"][0]["exp_avg"])
    # a state dict already in this implementation's layout passes through untouched
    assert AvatarModel.translate_optimizer_state(holder, ours.state_dict())["param_groups"][0]["params"] == [0]
