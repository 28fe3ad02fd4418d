"""CPU tier: the C-ABI library builds, loads without a GPU and exports every symbol ``include/sdt_b200.h`` declares;
the ctypes signature table covers exactly those symbols; product code never imports the oracle."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "sdt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(sdt_lib):
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(sdt_lib, n), f"{n} declared in include/sdt_b200.h but not exported"


def test_ctypes_table_matches_header():
    from scal_sdt_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_error_channel_without_gpu(sdt_lib):
    assert sdt_lib.sdt_version() >= 100
    # argument validation happens before any CUDA call, so it is observable on a CPU-only machine
    rc = sdt_lib.sdt_noise_target(None, None, None, None, 1000, None, None, 0, 1, 1, 0, None, None)
    assert rc == -1 and b"null pointer" in sdt_lib.sdt_last_error()
    assert sdt_lib.sdt_mse_loss_workspace_bytes() > 0
    assert sdt_lib.sdt_comm_world() == 0


def test_ops_fail_loudly_on_cpu_tensors():
    import torch
    from scal_sdt_b200 import NoiseScheduler, SdtError, get_lora
    m = get_lora(torch.nn.Linear(8, 8), rank=4)
    with pytest.raises(SdtError, match="no CPU path"):
        m(torch.randn(2, 8))
    with pytest.raises(SdtError, match="no CPU path"):
        NoiseScheduler().add_noise(torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8), torch.tensor([1]))


def test_product_code_never_imports_oracle():
    for path in (ROOT / "scal_sdt_b200").rglob("*.py"):
        src = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
