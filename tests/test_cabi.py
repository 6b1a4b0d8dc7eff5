"""The C-ABI library loads and exports every symbol include/cellbender_b200.h declares (no compute calls)."""
import os
import subprocess

import pytest

from cellbender_b200 import _cabi, build


@pytest.fixture(scope="module")
def built():
    build.build()
    return _cabi.lib()


def test_library_exports_every_declared_symbol(built):
    protos = _cabi.parse_header()
    assert len(protos) >= 24
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [n for n in protos if n not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_launch_count_table_covers_the_header():
    """gpu_launches (bench.py) is counted per C-ABI call from _cabi_counts: every declared entry point is either a
    launching call with a count or a host-side query, and the table names nothing the header does not declare."""
    from cellbender_b200._cabi_counts import LAUNCHES, QUERIES

    declared = set(_cabi.parse_header()) - {"cb_last_error"}
    assert set(LAUNCHES) | set(QUERIES) == declared, (sorted(declared - set(LAUNCHES) - set(QUERIES)),
                                                      sorted((set(LAUNCHES) | set(QUERIES)) - declared))
    assert not set(LAUNCHES) & set(QUERIES) and all(v >= 1 for v in LAUNCHES.values())


def test_abi_version_and_error_string(built):
    assert built.raw("cb_abi_version")() == 1
    assert isinstance(built.last_error(), str)


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.CabiError):
        _cabi._Lib()


def test_ops_refuse_cpu_tensors(built):
    import torch

    from cellbender_b200 import ops

    with pytest.raises(RuntimeError):
        ops.col_logsumexp(torch.zeros(4, 4), 4, 4)
