"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and the Python mirror keeps the reference's constructor / error behaviour.  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "cfa_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from clip_finegrained_alignment_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/cfa_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert _lib.lib.cfa_abi_version() == 4
    assert _lib.lib.cfa_error_string(-2).startswith(b"cfa:")
    assert _lib.lib.cfa_adamspd_chunk_elems() == 8192


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/cfa_b200.h must compile as C99 on its own (no C++ / CUDA / torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "h.c"
    src.write_text('#include "cfa_b200.h"\nint main(void) { return cfa_abi_version() == CFA_ABI_VERSION ? 0 : 1; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_descriptor_struct_matches_header():
    from clip_finegrained_alignment_b200.optimizers import _TENSOR_DT
    assert _TENSOR_DT.itemsize == 88
    assert [n for n in _TENSOR_DT.names][:7] == ["p", "g", "m", "v", "pre", "vmax", "numel"]


def test_sparc_capacity_query():
    from clip_finegrained_alignment_b200 import _lib
    assert _lib.lib.cfa_sparc_max_patches(77, 1) >= 257     # ViT-B/32, B/16 (+CLS), L/14@224 backward
    assert _lib.lib.cfa_sparc_max_patches(77, 0) >= 400


def test_adamspd_constructor_errors_match_reference():
    from clip_finegrained_alignment_b200 import AdamSPD
    p = torch.nn.Parameter(torch.zeros(4))
    for kw, msg in ((dict(lr=-1.0), "Invalid learning rate"), (dict(eps=-1.0), "Invalid epsilon"),
                    (dict(betas=(1.0, 0.9)), "index 0"), (dict(betas=(0.9, 1.0)), "index 1"),
                    (dict(weight_decay=-0.1), "Invalid weight_decay")):
        with pytest.raises(ValueError, match=msg):
            AdamSPD([p], **kw)
    opt = AdamSPD([{"params": [p], "pre": None}], lr=1e-3, amsgrad=True)
    assert opt.defaults["amsgrad"] is True and opt.param_groups[0]["pre"] is None
    assert opt.step() is None                      # no grads: nothing to do, no device needed
    assert opt.step(lambda: torch.tensor(3.0)) == 3.0


def test_state_dict_is_interchangeable_with_the_reference():
    """Checkpoint compatibility (finetuner.py:256-273 saves optimizer.state_dict()): a state_dict made
    by the reference AdamSPD loads into this one and back, including group['pre'] and the unused 'hyper' state."""
    from conftest import load_golden, reference_modules
    from clip_finegrained_alignment_b200 import AdamSPD
    f = load_golden("statedict_ref.pt")
    sd = f["state_dict"]
    params = [torch.nn.Parameter(x.clone()) for x in f["params"]]
    opt = AdamSPD([{"params": params, "pre": [torch.zeros_like(p) for p in params]}], lr=5.0, amsgrad=False)
    opt.load_state_dict(sd)
    grp = opt.param_groups[0]
    assert grp["lr"] == 1e-3 and grp["weight_decay"] == 0.1 and grp["amsgrad"] is True
    assert len(grp["pre"]) == 2 and torch.equal(grp["pre"][0], sd["param_groups"][0]["pre"][0])
    for j, p in enumerate(params):
        st = opt.state[p]
        assert st["step"] == 3 and isinstance(st["step"], int)
        assert set(st.keys()) == {"step", "exp_avg", "exp_avg_sq", "hyper", "max_exp_avg_sq"}
        assert torch.equal(st["exp_avg"], sd["state"][j]["exp_avg"])
    out = opt.state_dict()
    assert out["state"].keys() == sd["state"].keys() and out["param_groups"][0].keys() == sd["param_groups"][0].keys()
    mods = reference_modules()
    if mods is not None:                                  # and back into the unmodified reference optimizer
        ref = mods[1].AdamSPD([{"params": params, "pre": None}], lr=1.0)
        ref.load_state_dict(out)
        assert ref.state[params[1]]["step"] == 3 and ref.param_groups[0]["amsgrad"] is True


def test_no_cpu_fallback():
    from clip_finegrained_alignment_b200 import AdamSPD, CustomCLIPLoss, SPARCLoss, _lib
    import types
    cfg = types.SimpleNamespace(similarity_threshold=0.5, global_loss_weight=1.0, local_loss_weight=1.0,
                                inverse_temperature=1.0)
    with pytest.raises(_lib.CfaError):
        SPARCLoss(cfg)(torch.randn(2, 5, 8), torch.randn(2, 3, 8), torch.ones(2, 3, dtype=torch.bool))
    with pytest.raises(_lib.CfaError):
        CustomCLIPLoss()(torch.randn(4, 8), torch.randn(4, 8))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(_lib.CfaError):
        AdamSPD([{"params": [p], "pre": None}]).step()
    with pytest.raises(KeyError):
        AdamSPD([p]).step()                         # group without 'pre' (optimizers.py:146)
    with pytest.raises(TypeError):
        SPARCLoss(cfg)(torch.randn(2, 5, 8), torch.randn(2, 3, 8), torch.ones(2, 3))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing in the package, bench product path or build may import it."""
    pkg = os.path.join(ROOT, "clip_finegrained_alignment_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_launch_table_holds_ints_for_declared_symbols():
    """_lib.LAUNCHES feeds bench.py's gpu_launches: every value is a launch count, every key a declared entry point."""
    from clip_finegrained_alignment_b200 import _lib
    for name, n in _lib.LAUNCHES.items():
        assert isinstance(n, int) and n >= 0, (name, n)
        assert name in _lib.SIGNATURES, name
