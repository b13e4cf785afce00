"""CPU: the C-ABI library loads and exports every symbol include/pixelpick_b200.h declares."""
import ctypes
import os
import re

from pixelpick_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "pixelpick_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _header_functions() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    l = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(l, name), name


def test_version_and_error_string():
    l = _lib.lib()
    assert l.pp_version() >= 100
    assert isinstance(l.pp_last_error(), bytes)


def test_argument_validation_without_gpu():
    # workspace query is pure host arithmetic; bad shapes are rejected before any CUDA call
    l = _lib.lib()
    sz = ctypes.c_size_t()
    assert l.pp_acq_topk_workspace_bytes(4, 256 * 512, 6553, ctypes.byref(sz)) == 0
    assert sz.value > 4 * 8192 * 8
    assert l.pp_acq_topk_workspace_bytes(4, 100, 200, ctypes.byref(sz)) == -1
    assert b"bad" in l.pp_last_error()
    assert l.pp_acq_topk_workspace_bytes(1, (1 << 22) + 4, 10, ctypes.byref(sz)) == -1


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pixelpick_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f


def test_every_entry_point_rejects_null_arguments_before_any_cuda_call():
    """Each int-returning entry point, called with NULL pointers and zero sizes, returns PP_ERR_INVALID_ARG and a message
    naming itself (or the entry point it forwards to) - no CUDA call, no crash, so it holds on a host without a GPU."""
    l = _lib.lib()
    skip = {"pp_device_info",            # queries the device first (PP_ERR_CUDA here)
            "pp_acq_session_destroy",    # destroy(NULL) is a no-op, like free(NULL)
            "pp_conv_set_epilogue"}      # a process-wide switch: takes no pointer, returns the previous setting
    checked = 0
    for name, (argtypes, restype) in sorted(_lib._SIGNATURES.items()):
        if restype is not ctypes.c_int or not argtypes or name in skip:
            continue
        args = []
        for a in argtypes:
            if a in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(a, type) and issubclass(a, ctypes._Pointer)):
                args.append(None)
            else:
                args.append(0.0 if a is ctypes.c_float else 0)
        assert getattr(l, name)(*args) == -1, name
        assert l.pp_last_error().startswith(b"pp_"), (name, l.pp_last_error())
        checked += 1
    assert checked >= 35
    assert l.pp_acq_session_destroy(None) == 0


def test_integration_doc_covers_every_entry_point():
    """INTEGRATION.md maps each C entry point to the reference call site it replaces (families are written `pp_x(_suffix)`
    or `pp_x*` / `pp_x_*`)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stems = set(re.findall(r"pp_[a-z0-9_]+", doc))
    families = [m[:-1] for m in re.findall(r"pp_[a-z0-9_]+\*", doc)]
    for base, suf in re.findall(r"`(pp_[a-z0-9_]+)\((_[a-z0-9_]+)\)`", doc):
        stems.add(base + suf)
    for short in re.findall(r"`(_[a-z0-9_]+)`", doc):          # "`pp_dwconv3x3_fwd`, `_fwd_bnact`, `_dgrad`"
        stems.update(s.rsplit("_", 1)[0] + short for s in list(stems) if s.count("_") >= 2)
        stems.update(re.sub(r"_[a-z0-9]+$", "", s) + short for s in list(stems))
    missing = [n for n in _header_functions() if n not in stems and not any(n.startswith(f) for f in families)]
    assert not missing, missing
