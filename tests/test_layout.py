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


def test_running_mean_decorator_on_cpu_tensors():
    """utils.agg.mean_aggregator (src/utils/agg.py:6-91): value passthrough, masked / unmasked accumulation, scalars, reset."""
    import torch
    from semanticlidarunc_b200.utils.agg import mean_aggregator

    @mean_aggregator()
    def twice(x):
        """doc"""
        return 2 * x

    assert twice.__name__ == "twice" and twice.__doc__ == "doc"
    a = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    assert torch.equal(twice(a), 2 * a) and twice.mean() == 0.0
    out = twice.accumulate(a)
    assert torch.equal(out, 2 * a) and abs(twice.mean() - 5.0) < 1e-12
    twice.accumulate(a, mask=torch.tensor([[True, False, False], [False, False, True]]))     # adds 0 and 10 over 2 elements
    assert abs(twice.mean() - (30.0 + 10.0) / 8) < 1e-12
    twice.add(4.0)
    assert abs(twice.mean(reset=True) - 44.0 / 9) < 1e-12 and twice.mean() == 0.0
    twice.sync_ddp()                                         # no process group: a no-op


def test_running_mean_ignores_non_finite_values_outside_the_mask():
    """src/utils/agg.py:52 sums x[mask]: NaN / Inf at masked-out pixels must not poison the running mean."""
    import torch
    from semanticlidarunc_b200.utils.agg import mean_aggregator

    @mean_aggregator()
    def ident(x):
        return x

    x = torch.tensor([1.0, float("nan"), 3.0, float("inf")])
    ident.accumulate(x, mask=torch.tensor([True, False, True, False]))
    assert ident.mean() == 2.0
