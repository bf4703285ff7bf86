"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = ""
    for h in ("rvb_b200.h", "rvb_diar.h"):
        with open(os.path.join(ROOT, "include", h)) as f:
            text += f.read()
    return sorted(set(re.findall(r"RVB_API[^;(]*?\b(rvb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from reverb_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by librvb_b200.so"
    assert sorted(_lib.SIGNATURES) == names, "reverb_b200/_lib.py must bind exactly the header's symbols"


def test_pure_host_entry_points():
    from reverb_b200 import _lib
    lib = _lib.load()
    assert lib.rvb_fbank_num_frames(399) == 0
    assert lib.rvb_fbank_num_frames(400) == 1
    assert lib.rvb_fbank_num_frames(480000) == 2998
    assert lib.rvb_encoder_out_frames(2998) == 748 and lib.rvb_encoder_out_frames(2051) == 512
    assert lib.rvb_launch_count() == 0 or lib.rvb_launch_count() > 0


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        return
    from reverb_b200 import _lib
    from reverb_b200._lib import ModelConfig
    lib = _lib.load()
    cfg = ModelConfig()
    h = lib.rvb_model_create(ctypes.byref(cfg))
    assert not h
    assert "no CUDA device" in _lib.last_error()


def test_product_never_imports_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may touch oracle/ — the package itself must not."""
    pkg = os.path.join(ROOT, "reverb_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_ctypes_struct_mirrors_match_the_c_headers(tmp_path):
    """sizeof of every config struct as gcc sees it in include/*.h == ctypes.sizeof of its mirror in reverb_b200/_lib.py
    (the headers are plain C: they must compile with a C compiler, without CUDA)."""
    import shutil
    import subprocess
    from reverb_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rvb_diar.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu\\n", sizeof(rvb_model_config), sizeof(rvb_seg_config), sizeof(rvb_emb_config));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / "sz"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.ModelConfig), ctypes.sizeof(_lib.SegConfig), ctypes.sizeof(_lib.EmbConfig)]
