"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/pnce.h
declares; host-only entry points behave; the product path has no CPU fallback and never imports
the oracle."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gan_variant_research_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from gan_variant_research_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pnce.h")).read()
    declared = sorted(set(re.findall(r"\b(pnce_[a-z_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pnce.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    # ... and nothing else leaves the library: no experiment hooks (pnce_debug_*), no kernel launch stubs
    import subprocess
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = sorted(line.split()[-1] for line in nm.stdout.splitlines() if line.split()[1:2] and line.split()[1] in "TDBRW")
    assert exported == declared, sorted(set(exported) ^ set(declared))
    assert lib.pnce_abi_version() == _lib.ABI_VERSION


def test_ctypes_struct_matches_header_layout():
    from gan_variant_research_b200 import _lib
    assert ctypes.sizeof(_lib.PnceLayer) == 4 * 8 + 4 * 4
    assert _lib.PnceLayer.C.offset == 32 and _lib.PnceLayer.P.offset == 44


def _layers(shapes, p):
    from gan_variant_research_b200 import _lib
    arr = (_lib.PnceLayer * len(shapes))()
    for i, (c, h, w) in enumerate(shapes):
        arr[i].C, arr[i].H, arr[i].W, arr[i].P = c, h, w, min(p, h * w)
    return arr


def test_workspace_query_and_argument_errors(lib):
    n = ctypes.c_size_t(0)
    b5 = [(64, 256, 256), (256, 64, 64), (256, 64, 64), (128, 128, 128), (64, 256, 256)]
    assert lib.pnce_workspace_bytes(_layers(b5, 256), 5, 64, ctypes.byref(n)) == 0
    rows = 64 * 256 * (64 + 256 + 256 + 128 + 64) * 4
    # dxT + (SIMT: 2 fp32 row sets | tensor-core: 6 bf16 operand blobs (q, k, key-major k; hi+lo) + raw fp32 q)
    assert 3 * rows <= n.value <= 5 * rows + (4 << 20)
    assert lib.pnce_workspace_bytes(_layers(b5, 256), 5, 0, ctypes.byref(n)) == -1       # batch < 1
    assert lib.pnce_workspace_bytes(_layers(b5, 256), 9, 1, ctypes.byref(n)) == -2       # > 8 layers
    assert lib.pnce_workspace_bytes(_layers([(2048, 4, 4)], 16), 1, 1, ctypes.byref(n)) == -2
    assert lib.pnce_workspace_bytes(_layers([(8, 128, 128)], 8192), 1, 1, ctypes.byref(n)) == -2
    assert lib.pnce_workspace_bytes(None, 1, 1, ctypes.byref(n)) == -1
    assert lib.pnce_status_string(-3).decode().startswith("workspace")
    assert lib.pnce_sample_bwd_workspace_bytes(2, 16, 8, 8, 32, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.pnce_rows_loss_workspace_bytes(2, 256, 256, ctypes.byref(n)) == 0 and n.value > 0


def test_compute_entry_points_reject_bad_arguments_without_touching_cuda(lib):
    lay = _layers([(8, 4, 4)], 16)
    # NULL workspace / loss pointers are rejected before any launch
    assert lib.pnce_fwd(lay, 1, 1, 0, 0.07, 0, None, 0, None, None, None) == -1
    assert lib.pnce_fwd(lay, 1, 1, 7, 0.07, 0, None, 0, None, None, None) == -1          # bad dtype
    assert lib.pnce_bwd(lay, 1, 1, 0, 0, None, 0, None, None) == -1                         # dtgt NULL


def test_no_cpu_fallback():
    import gan_variant_research_b200 as pn
    src = [torch.randn(1, 4, 4, 4)]
    tgt = [torch.randn(1, 4, 4, 4, requires_grad=True)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pn.PatchNCELoss(0.07, 8)(src, tgt)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pn.PatchSampleF()(tgt, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pn.rows_patchnce(torch.randn(8, 4), torch.randn(8, 4), num_patches=8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gan_variant_research_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "/root/reference" not in text, f


def test_reference_shim_replaces_the_module_path():
    import gan_variant_research_b200 as pn
    saved = sys.modules.get("GAN_Variant1.losses.patchnce_cut")
    try:
        mod = pn.install_reference_shim()
        import importlib
        got = importlib.import_module("GAN_Variant1.losses.patchnce_cut")
        assert got is mod
        assert got.compute_patchnce_loss is pn.compute_patchnce_loss
        assert got.PatchNCELoss is pn.PatchNCELoss
    finally:
        if saved is not None:
            sys.modules["GAN_Variant1.losses.patchnce_cut"] = saved
        else:
            sys.modules.pop("GAN_Variant1.losses.patchnce_cut", None)


def test_patch_count_and_signature_defaults():
    import inspect
    import gan_variant_research_b200 as pn
    assert pn.patch_count(256, 100) == 100 and pn.patch_count(256, 65536) == 256
    sig = inspect.signature(pn.compute_patchnce_loss)
    assert list(sig.parameters)[:6] == ["generator", "src_images", "tgt_images", "nce_layers",
                                        "temperature", "num_patches"]
    assert sig.parameters["temperature"].default == 0.07 and sig.parameters["num_patches"].default == 256
    m = pn.PatchNCELoss()
    assert (m.temperature, m.num_patches, m.nce_layers) == (0.07, 256, [0, 4, 8, 12, 16])
    s = inspect.signature(pn.PatchSampleF.forward)
    assert list(s.parameters)[1:] == ["feats", "num_patches", "patch_ids"]


def test_id_draw_helper_follows_the_reference_order_on_cpu():
    """draw_patch_ids_all = one randint per layer in layer order (patchnce_cut.py:36-38, :63); on CPU
    tensors (no side stream) it must consume the global generator exactly like the reference loop."""
    import gan_variant_research_b200 as pn
    shapes = [(8, 16, 16), (4, 3, 3), (16, 32, 32)]
    feats = [torch.zeros(2, *s) for s in shapes]
    torch.manual_seed(7)
    got = pn.draw_patch_ids_all(feats, 64)
    after_got = torch.rand(3)
    torch.manual_seed(7)
    want = [torch.randint(0, s[1] * s[2], (min(64, s[1] * s[2]),)) for s in shapes]
    after_want = torch.rand(3)
    assert all(torch.equal(g, w) and g.dtype == torch.int64 for g, w in zip(got, want))
    assert torch.equal(after_got, after_want)
    assert pn.draw_patch_ids_all([], 64) == []


def test_pinned_view_refuses_pageable_memory():
    import gan_variant_research_b200 as pn
    with pytest.raises(RuntimeError, match="pinned"):
        pn.pinned_as_device(torch.zeros(4, 4))


def test_every_kernel_launched_with_the_pdl_attribute_waits_for_its_predecessors(lib):
    """Programmatic dependent launch is only safe because every kernel that may carry the launch attribute
    executes griddepcontrol.wait (SASS: ACQBULK) before its first global access: a kernel added without pdl_enter() but
    launched through launch_k's default site would race with its predecessor silently.  Checked on the built library:
    the kernels WITHOUT the wait are exactly the dense-backward ones, and the host launches exactly those with
    launch_k<0> (never the attribute)."""
    import re
    import shutil
    import subprocess
    from gan_variant_research_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not found")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    waits, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            waits[name] = False
        elif name is not None and "ACQBULK" in line:
            waits[name] = True
    assert len(waits) >= 40, "no kernels found in the library's SASS"
    no_wait = {re.search(r"\d+(k_[a-z_]+?)(?:I|E)", n).group(1) for n, w in waits.items() if not w}     # mangled -> k_name
    dense = {"k_dense_flat", "k_dense_direct", "k_dense_nhwc", "k_fill_zero", "k_scatter_nhwc"}
    assert no_wait == dense, no_wait ^ dense
    src = open(os.path.join(ROOT, "gan_variant_research_b200", "csrc", "pnce_api.cu")).read()
    site0 = set(re.findall(r"launch_k<0>\((k_[a-z_]+)", src))
    assert site0 == dense, site0 ^ dense
    for k in dense:                                              # and never through another site or a bare <<< >>>
        assert not re.search(r"launch_k(?:<[1-9]\d*>)?\(" + k + r"\b", src), k
        assert not re.search(k + r"(?:<[^<>;]*>)?<<<", src), k
