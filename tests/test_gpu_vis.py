"""GPU (-m gpu): row f3, the explanation output path -- ``scouter_vis_upsample_u8`` against Pillow's own outputs
(tests/golden/vis_upsample.npz) and the oracle restatement (oracle/vis.py), bit-exact (byte work), and
``SlotModel.explain`` end to end."""
import numpy as np
import pytest
import torch

from conftest import load_golden, vis_cases
from oracle import vis as ov
from oracle.refshim import make_args
import scouter_b200 as sb
from scouter_b200 import _lib as L
from scouter_b200.synth import fill_state_dict, synth_images

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    L.check(L.lib().scouter_device_check(0))
    return torch.device("cuda", 0)


def upsample(maps_np, oh, ow, dev, want_ratios=True):
    maps = torch.from_numpy(np.ascontiguousarray(maps_np)).to(dev)
    c, h, w = maps.shape
    out = torch.zeros(c, oh, ow, dtype=torch.uint8, device=dev)
    ratios = torch.zeros(c, dtype=torch.float64, device=dev) if want_ratios else None
    L.check(L.lib().scouter_vis_upsample_u8(maps.data_ptr(), c, h, w, oh, ow, out.data_ptr(), L.ptr(ratios), 0),
            "scouter_vis_upsample_u8")
    torch.cuda.synchronize()
    return out.cpu().numpy(), (ratios.cpu().numpy() if want_ratios else None)


def test_upsample_vs_pillow_golden(dev):
    cases, _ = vis_cases()
    for name, (maps, heat, ratios) in cases.items():
        got, r = upsample(maps, heat.shape[1], heat.shape[2], dev)
        assert np.array_equal(got, heat), name
        assert np.array_equal(r, ratios), name


@pytest.mark.parametrize("c,h,w,oh,ow", [(10, 7, 7, 224, 224), (30, 9, 9, 260, 260), (200, 7, 7, 224, 224), (3, 7, 9, 481, 1023),
                                         (2, 9, 9, 1, 1), (2, 31, 17, 8, 5), (1, 3, 2, 5, 3000)])
def test_upsample_vs_oracle_fresh_maps(dev, c, h, w, oh, ow):
    maps = np.random.RandomState(c * 1000 + oh).randint(0, 256, (c, h, w)).astype(np.uint8)
    got, r = upsample(maps, oh, ow, dev)
    assert np.array_equal(got, ov.resize_bilinear_u8(maps, oh, ow))
    assert np.array_equal(r, np.array([ov.attention_ratio(m) for m in maps]))


def test_upsample_properties_and_errors(dev):
    # constant maps stay constant; same-size resize is the identity; ratios alone (out = NULL) work
    const = np.full((2, 7, 7), 200, np.uint8)
    assert (upsample(const, 224, 224, dev)[0] == 200).all()
    maps = np.random.RandomState(5).randint(0, 256, (4, 9, 9)).astype(np.uint8)
    assert np.array_equal(upsample(maps, 9, 9, dev)[0], maps)
    m = torch.from_numpy(maps).to(dev)
    ratios = torch.zeros(4, dtype=torch.float64, device=dev)
    L.check(L.lib().scouter_vis_upsample_u8(m.data_ptr(), 4, 9, 9, 0, 0, 0, ratios.data_ptr(), 0))
    torch.cuda.synchronize()
    assert np.array_equal(ratios.cpu().numpy(), np.array([ov.attention_ratio(x) for x in maps]))
    assert L.lib().scouter_vis_upsample_u8(0, 4, 9, 9, 8, 8, m.data_ptr(), 0, 0) == -1          # NULL maps
    assert L.lib().scouter_vis_upsample_u8(m.data_ptr(), 4, 9, 9, 0, 8, m.data_ptr(), 0, 0) == -1  # bad output size
    with pytest.raises(sb.ScouterError):
        L.check(L.lib().scouter_vis_upsample_u8(m.data_ptr(), 0, 9, 9, 8, 8, m.data_ptr(), 0, 0))


def test_explain_end_to_end(dev):
    """SlotModel.explain = forward + vis branch (slot_attention.py:68-80) + test.py:35,43, all on the device: the maps
    equal the module's own vis output, and the heat maps / ratios equal the oracle applied to those maps."""
    z, meta = load_golden("cfg2_resnest26d_pos_224")
    m = sb.SlotModel(make_args(**meta["args"]))
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
    m = m.to(dev).eval()
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"]).to(dev)
    with torch.no_grad():
        plain = m(x)
        e = m.explain(x, vis_id=1)
        e2 = m.explain(x, vis_id=1, out_size=(100, 150))
    assert torch.equal(e["log_probs"], plain)
    maps = e["maps"].cpu().numpy()
    assert maps.shape == (10, 7, 7) and maps.dtype == np.uint8 and maps.min() == 0 and maps.max() == 255
    assert e["heat"].shape == (10, 224, 224) and e2["heat"].shape == (10, 100, 150)
    assert np.array_equal(e["heat"].cpu().numpy(), ov.resize_bilinear_u8(maps, 224, 224))
    assert np.array_equal(e2["heat"].cpu().numpy(), ov.resize_bilinear_u8(maps, 100, 150))
    assert np.array_equal(e["ratios"].cpu().numpy(), np.array([ov.attention_ratio(v) for v in maps]))
    assert not m.keep_attn                                   # explain() restores the flag
    with pytest.raises(sb.ScouterError):
        m.explain(x, vis_id=meta["batch"])
