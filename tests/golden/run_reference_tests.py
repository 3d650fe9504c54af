"""Run the reference's OWN test-suite on the NumPy stand-in for JAX (build container only: needs /root/reference).

    python tests/golden/run_reference_tests.py

`tests/core/*.py` check the SparseTensor products against einsum, the elemental rules and shape-op transforms
against `jax.jacrev`; `tests/examples/example_test.py` checks `graphax.jacve` on every example (forward, reverse and
custom orders, per-sample and vmapped) against `jax.jacrev` with the reference's own `tree_allclose`.  On the
stand-in `jax.jacrev` is torch.autograd in float64 on the traced equations (`jaxlite/jax/_autodiff.py`), i.e. an
independent autodiff.  Writes `tests/golden/reference_tests_on_jaxlite.json`.
"""
import json
import re
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
FILES = ["tests/core/sparse_mul_test.py", "tests/core/dense_mul_test.py", "tests/core/mixed_mul_test.py",
         "tests/core/replication_test.py", "tests/core/primitive_test.py", "tests/core/multi_output_primitives_test.py",
         "tests/examples/example_test.py", "tests/examples/hessian_test.py"]
KNOWN = {
    "PrimitveTest::test_squeezing": "reference: primitives.py:927 calls list.index(int) on a list of Dimension objects (ValueError)",
    "MultiOutputTest::test_split": "stand-in: make_jaxpr THROUGH jacve (printing the emitted program) needs traceable index arrays",
    "ExampleTests::test_Encoder": "the custom 93-vertex order is written for the jaxpr of the author's JAX version "
                                  "(jax.nn.softmax lowers differently); its fwd / rev halves are covered by reference_run",
}
out = {"reference": str(REF), "files": {}}
env = {"PYTHONPATH": f"{HERE / 'jaxlite'}:{REF / 'src'}", "PATH": "/usr/bin:/bin:" + str(Path(sys.executable).parent)}
for f in FILES:
    r = subprocess.run([sys.executable, "-m", "pytest", "-c", "/dev/null", "-p", "no:cacheprovider", "--rootdir=/tmp", "-q",
                        "-rA", str(REF / f)], capture_output=True, text=True, env=env, cwd="/tmp")
    passed = sorted(set(re.findall(r"^PASSED \S*?::(\S+::\S+)", r.stdout, re.M)))
    failed = sorted(set(re.findall(r"^FAILED \S*?::(\S+::\S+)", r.stdout, re.M)))
    # example_test.py defines helpers named test_order / test_fwd / test_rev that pytest collects: not tests
    out["files"][f] = {"passed": len(passed), "failed": {t: KNOWN.get(t, "unexplained") for t in failed}}
    print(f, len(passed), "passed", failed)
out["total_passed"] = sum(v["passed"] for v in out["files"].values())
out["total_failed"] = sum(len(v["failed"]) for v in out["files"].values())
(HERE / "reference_tests_on_jaxlite.json").write_text(json.dumps(out, indent=1))
print(out["total_passed"], "passed,", out["total_failed"], "failed")
