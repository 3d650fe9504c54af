"""The f32x2 pairing pass (graphax_b200/slp.py, opt-in GX_SLP=1): host-side invariants, no GPU needed.

Every plan operation is emitted exactly once and after its operands; the two halves of a packed instruction are
independent operations of one kind; and - because ptxas contracts ``mul.rn.f32x2`` feeding ``add.rn.f32x2`` into FFMA2
(tools/microbench/f32x2_contraction.cu) - no packed sum reads the pair a packed product wrote."""
import pytest

from graphax_b200 import codegen
from graphax_b200.configs import get_workload
from graphax_b200.slp import pair_operations


@pytest.mark.parametrize("workload,order", [("roe3d", "mixed"), ("roe3d", "rev"), ("roe1d", "rev"), ("readme", "readme")])
def test_pairing_is_a_valid_emission_order(workload, order):
    plan = get_workload(workload).plan(order, rounding="reference")
    pairing = pair_operations(plan)
    dag = plan.dag
    defined, emitted = set(), []
    for what, item in pairing.order:
        ops = [item] if what == "op" else list(pairing.packed[item][:2])
        if what == "pack":
            op0, op1, _ = pairing.packed[item]
            assert {op0.op, op1.op} <= {"mul"} or {op0.op, op1.op} <= {"add", "sub"}
            assert op0.dst not in op1.args and op1.dst not in op0.args
        for op in ops:
            for a in op.args:
                n = a
                while dag.op[n] == "neg":
                    n = dag.args[n][0]
                assert n in defined or dag.op[n] in ("const", "input") or not any(o.dst == n for o in plan.ssa), (op, a)
            if not op.op.startswith("store"):
                defined.add(op.dst)
            emitted.append(op)
    assert sorted(map(id, emitted)) == sorted(map(id, plan.ssa))        # every operation exactly once
    mul_packs = {(p[0].dst, p[1].dst) for p in pairing.packed if p[0].op == "mul"}
    for op0, op1, operands in pairing.packed:
        if op0.op == "mul":
            continue
        for a0, a1 in operands:
            while dag.op[a0] == "neg":
                a0 = dag.args[a0][0]
            while dag.op[a1] == "neg":
                a1 = dag.args[a1][0]
            assert (a0, a1) not in mul_packs


def test_generated_source_with_pairs(monkeypatch):
    plan = get_workload("roe3d").plan("rev", rounding="reference")
    plain = codegen.emit_source(plan, nvrtc=True)
    monkeypatch.setenv("GX_SLP", "1")
    paired = codegen.emit_source(plan, nvrtc=True)
    assert "__fmul2_rn" in paired and "__fadd2_rn" in paired and paired != plain
    monkeypatch.setenv("GX_SLP", "0")
    assert codegen.emit_source(plan, nvrtc=True) == plain


def test_two_sample_body_keeps_the_reference_rounding(tmp_path):
    """samples_per_thread=2 packs both samples' operations into f32x2 instructions; sums that read a product stay
    scalar, so that ptxas cannot contract them: the compiled kernel must not contain a single FFMA2."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    plan = get_workload("roe1d").plan("rev", rounding="reference")
    src = codegen.emit_source(plan, nvrtc=False, samples_per_thread=2, threads=128, min_blocks=2)
    assert "make_float2(__fadd_rn(" in src and "__fmul2_rn(" in src
    cu = tmp_path / "k.cu"
    cu.write_text(src)
    cubin = tmp_path / "k.cubin"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-I", str(codegen.CSRC), "-cubin", "-o",
                    str(cubin), str(cu)], check=True, capture_output=True)
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], check=True, capture_output=True, text=True).stdout
    assert "FMUL2" in sass and "FFMA2" not in sass
