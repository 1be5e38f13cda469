"""CPU-side checks of the `net.CIDNet.CIDNet` module surface the reference's callers rely on (no compute):
PyTorchModelHubMixin save/from_pretrained (app.py:86, README), the safetensors `strict=False` load of eval_hf.py:21-35,
copy.deepcopy / pickle / torch.save of the module (EMA / DataLoader patterns), the 191-key state_dict surface."""
import copy
import io
import os
import pickle

import pytest
import torch

from conftest import GOLDEN


def _model(perturb=0.01):
    import hvi_cidnet_b200  # noqa: F401
    from hvi_cidnet_b200.net.CIDNet import CIDNet
    torch.manual_seed(3)
    m = CIDNet()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(perturb * torch.randn_like(p))
    return m, CIDNet


def test_state_dict_surface_matches_reference_keys():
    m, _ = _model()
    lines = open(os.path.join(GOLDEN, "state_dict_keys.txt")).read().strip().splitlines()
    ref = {l.split(" ", 1)[0]: eval(l.split(" ", 1)[1]) for l in lines}
    sd = m.state_dict()
    assert set(sd.keys()) == set(ref.keys()) and len(sd) == 191          # order is irrelevant for dict loading
    assert all(list(v.shape) == ref[k] and v.dtype == torch.float32 for k, v in sd.items())
    assert sum(v.numel() for v in sd.values()) == 1975569


def test_hub_mixin_save_and_from_pretrained_round_trip(tmp_path):
    """save_pretrained -> config.json + model.safetensors; from_pretrained(local dir) restores every tensor (app.py:86)."""
    m, CIDNet = _model()
    m.save_pretrained(str(tmp_path))
    assert {"config.json", "model.safetensors"} <= set(os.listdir(tmp_path))
    m2 = CIDNet.from_pretrained(str(tmp_path))
    for (k, a), (k2, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k


def test_safetensors_strict_false_load_like_eval_hf(tmp_path):
    """eval_hf.py:21-35: sf.load_file(model.safetensors) -> model.load_state_dict(sd, strict=False), here with a
    checkpoint that lacks one key and carries an unknown one (what strict=False is for)."""
    import safetensors.torch as sf
    m, CIDNet = _model()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    del sd["I_LCA5.ffn.temperature"]                       # dead block in the base graph
    sd["extra.unused"] = torch.zeros(3)
    path = os.path.join(tmp_path, "model.safetensors")
    sf.save_file(sd, path)
    fresh = CIDNet()
    res = fresh.load_state_dict(sf.load_file(path), strict=False)
    assert res.missing_keys == ["I_LCA5.ffn.temperature"] and res.unexpected_keys == ["extra.unused"]
    assert torch.equal(fresh.state_dict()["HV_LCA3.gdfn.project_in.weight"], sd["HV_LCA3.gdfn.project_in.weight"])
    with pytest.raises(RuntimeError):
        CIDNet().load_state_dict(sf.load_file(path), strict=True)


def test_deepcopy_pickle_and_torch_save_of_the_module():
    """the native handle / workspaces are not part of the module's state: copies start without a context"""
    m, CIDNet = _model()
    m.__dict__["_ctx"] = None
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone._ctx is None and clone._workspaces == {}
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), clone.state_dict().values()))
        assert clone.trans.density_k.data_ptr() != m.trans.density_k.data_ptr()
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert len(m3.state_dict()) == 191


def test_weight_signature_tracks_updates():
    m, _ = _model()
    s0 = m._weights_signature()
    assert s0 == m._weights_signature()
    with torch.no_grad():
        m.HV_LCA2.norm.weight.mul_(1.5)                    # in-place update bumps _version
    s1 = m._weights_signature()
    assert s1 != s0
    with torch.no_grad():
        m.trans.density_k.fill_(0.3)                       # read from device memory every call: not part of the signature
    assert m._weights_signature() == s1
    m.double()
    assert m._weights_signature() != s1                    # storage replaced
