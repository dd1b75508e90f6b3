"""CPU-only: libsac_b200.so loads and exports every symbol include/sacb.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sacb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sacb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from da_sac_b200 import lib
    assert os.path.isfile(lib.LIB_PATH), "run __graft_entry__.build() first"
    dll = ctypes.CDLL(lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(dll, s)]
    assert not missing, missing
    assert lib.lib().sacb_abi_version() == lib.ABI_VERSION
    assert lib.lib().sacb_aspp_jpad() == 768


def test_struct_sizes_match_header_layout():
    from da_sac_b200 import lib
    # natural alignment on LP64: 14 x int32 then pointers
    assert ctypes.sizeof(lib.ConvGemm) == 14 * 4 + 10 * 8 + 8 + 5 * 8 + 8      # ... colsum, precision (+ tail padding)
    assert ctypes.sizeof(lib.ConvWgrad) == 14 * 4 + 5 * 8 + 8


def test_ctypes_structs_match_the_header_as_gcc_lays_it_out(tmp_path):
    """sizeof + offset of the last field of every descriptor struct: include/sacb.h compiled by gcc vs the ctypes mirrors"""
    import subprocess
    from da_sac_b200 import lib, p2p
    pairs = [("SacbConvGemm", lib.ConvGemm, "precision"), ("SacbConvWgrad", lib.ConvWgrad, "precision"), ("SacbTail", lib.Tail, "pool_mode"),
             ("SacbLoss", lib.Loss, "grad_rows"), ("SacbAllreduceSgd", p2p.AllreduceSgd, "mc_params")]
    src = tmp_path / "sz.c"
    body = "".join('  printf("%%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));\n' % (c, c, last) for c, _, last in pairs)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sacb.h"\nint main(void) {\n' + body + "  return 0;\n}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    for (cname, ct, last), line in zip(pairs, out):
        size, off = [int(v) for v in line.split()]
        assert ctypes.sizeof(ct) == size, (cname, ctypes.sizeof(ct), size)
        assert getattr(ct, last).offset == off, (cname, last)


def test_missing_library_fails_loudly(monkeypatch):
    from da_sac_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libsac_b200.so")
    import pytest
    with pytest.raises(lib.SacbError):
        lib.lib()


def test_model_surface_matches_reference_contract():
    """state_dict keys / optimiser groups of the drop-in, no GPU needed"""
    import torch
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    net = get_model(synth.ModelCfg(), 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    sd = net.state_dict()
    assert len(sd) == 2 + 2 * 632
    assert "backbone.model.layer3.22.bn3.running_var" in sd and "slow_net.model.layer5.conv2d_list.3.bias" in sd
    assert sd["running_conf"].shape == (19,) and sd["slow_init"].shape == (1,)
    assert [len(g["params"]) for g in net.parameter_groups(2.5e-4, 5e-4)] == [208, 104, 4, 4]
    assert all(not p.requires_grad for p in net.slow_net.parameters())
    bb = synth.make_backbone_params()
    assert list(bb.keys()) == [k[len("backbone."):] for k in sd if k.startswith("backbone.")]
