"""oracle/build_ref_ext.py: the recipe that compiles the reference's own CUDA extension as a reported baseline.  Without the
prebuilt oracle/_ref/quaternion_ops.so the loader answers None (and bench.py then omits `reference_cuda_ext`); with it, the
module exposes the three entry points of quaternion_ops_py.cpp:132-165 — the same names our drop-in shim exports."""
from oracle import build_ref_ext


def test_loader_is_optional_and_names_match_the_shim():
    from quan_ultralytics_b200 import quaternion_ops as shim
    names = ("qconv_forward", "qconv_backward", "iqbn_forward")
    assert all(callable(getattr(shim, n)) for n in names)
    ref = build_ref_ext.load()
    if ref is not None:
        assert all(callable(getattr(ref, n)) for n in names)


def test_recipe_reads_sources_in_place_and_writes_only_under_oracle_ref():
    assert str(build_ref_ext.REF).startswith("/root/reference/")
    assert build_ref_ext.OUT == build_ref_ext.ROOT / "oracle" / "_ref"
    gitignore = (build_ref_ext.ROOT / ".gitignore").read_text().split()
    assert "oracle/_ref/" in gitignore
