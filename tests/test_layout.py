"""Repository rules: the product never imports the oracle, the reference tree is never read at
run time by the GPU tests / bench / smoke, and no reference source is copied in."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for dp, _, fs in os.walk(os.path.join(ROOT, d)):
        if "__pycache__" in dp or os.sep + "build" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(dp, f)


def test_product_package_never_touches_oracle_or_reference():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference", re.M)
    for f in _py_files("semanticlidarunc_b200"):
        assert not pat.search(open(f).read()), f


def test_gpu_tests_and_bench_do_not_read_reference_tree():
    for f in ["bench.py", "__graft_entry__.py"] + [os.path.join("tests", n) for n in os.listdir(os.path.join(ROOT, "tests"))
                                                    if n.startswith("test_gpu")]:
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert "/root/reference" not in open(p).read() and "_refshim" not in open(p).read(), f


def test_oracle_headers_say_test_infrastructure():
    for f in _py_files("oracle"):
        assert "TEST INFRASTRUCTURE ONLY" in open(f).read() or f.endswith("_refshim.py"), f
