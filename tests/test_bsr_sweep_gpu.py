"""GPU parity (pytest -m gpu) of the BSR sweep executor (csrc/bsr_sweep.cu) through the C ABI: C = A_bsr B for 3x3 blocks
and 64 columns against the oracle's BSR product (oracle/oracle_spmv.c: oracle_bsr_spmm, the restatement of the CitcomS
node operator citcoms/lib/Element_calculations.c:516-571), tolerance 1e-12 * (|A| |B|)."""
import numpy as np
import pytest
import scipy.sparse as sp

from test_bsr_plan_cpu import stencil_bsr

pytestmark = pytest.mark.gpu


def run_plan(rp, ci, blocks, Bd, strips):
    import torch

    from g4s_b200 import bsr

    mb = len(rp) - 1
    t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rp, ci, blocks.reshape(-1), Bd.reshape(-1))]
    plan = bsr.BsrPlan(mb, Bd.shape[0] // 3, t[0].data_ptr(), t[1].data_ptr() if len(ci) else 0, strips)
    plan.set_values(t[2].data_ptr() if len(ci) else 0)
    out = torch.full((mb * 3 * 64,), -7.0, dtype=torch.float64, device="cuda")
    plan.spmm(t[3].data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    return plan, t, out


def compare(oracle, rp, ci, blocks, Bd, out):
    want = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), 3, Bd)
    absw = oracle.bsr_spmm(rp, ci, np.abs(blocks).reshape(-1), 3, np.abs(Bd))
    got = out.cpu().numpy().reshape(want.shape)
    assert np.all(np.abs(got - want) <= 1e-12 * absw + 1e-300)


@pytest.mark.parametrize("shape", [(9, 6, 7), (4, 4, 40), (17, 13, 5), (1, 1, 1)])
def test_sweep_on_mesh_operator_matches_oracle(oracle, shape):
    from g4s_b200 import bsr

    n0, n1, n2 = shape
    rng = np.random.default_rng(n0 * 100 + n1)
    rp, ci, blocks = stencil_bsr(n0, n1, n2, rng)
    Bd = rng.uniform(-1, 1, (n0 * n1 * n2 * 3, 64))
    plan, t, out = run_plan(rp, ci, blocks, Bd, bsr.grid_pencil_strips(n0, n1, 0, n2))
    compare(oracle, rp, ci, blocks, Bd, out)
    if n0 * n1 * n2 > 100:
        assert plan.info()["slot_fill"] > 0.5
    # new values, same pattern: repack and multiply again (what a nonlinear solver does every outer iteration)
    import torch

    blocks2 = rng.uniform(-1, 1, blocks.shape)
    v2 = torch.from_numpy(blocks2.reshape(-1)).cuda()
    plan.set_values(v2.data_ptr())
    plan.spmm(t[3].data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    compare(oracle, rp, ci, blocks2, Bd, out)


@pytest.mark.parametrize("seed", [0, 1])
def test_sweep_on_unstructured_pattern_default_strips(oracle, seed):
    rng = np.random.default_rng(seed)
    mb = 1000 + 37 * seed
    pat = (sp.random(mb, mb, density=0.01, random_state=rng, format="csr") + sp.diags([1.0] * 3, [-1, 0, 1], (mb, mb))).tocsr()
    pat.sort_indices()
    rp, ci = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
    blocks = rng.uniform(-1, 1, (pat.nnz, 3, 3))
    Bd = rng.uniform(-1, 1, (mb * 3, 64))
    _, _, out = run_plan(rp, ci, blocks, Bd, None)
    compare(oracle, rp, ci, blocks, Bd, out)


def test_sweep_empty_rows_and_empty_matrix(oracle):
    rp = np.array([0, 0, 2, 2, 3], dtype=np.int32)
    ci = np.array([3, 0, 2], dtype=np.int32)
    blocks = np.random.default_rng(1).uniform(-1, 1, (3, 3, 3))
    Bd = np.random.default_rng(2).uniform(-1, 1, (12, 64))
    _, _, out = run_plan(rp, ci, blocks, Bd, None)
    compare(oracle, rp, ci, blocks, Bd, out)
    assert np.all(out.cpu().numpy().reshape(12, 64)[0:3] == 0.0)  # an empty row is written as zeros, not left untouched


def test_spmm_before_set_values_is_an_error():
    import torch

    from g4s_b200 import G4SError, bsr

    rp = torch.tensor([0, 1], dtype=torch.int32, device="cuda")
    ci = torch.tensor([0], dtype=torch.int32, device="cuda")
    plan = bsr.BsrPlan(1, 1, rp.data_ptr(), ci.data_ptr())
    B = torch.zeros(192, dtype=torch.float64, device="cuda")
    with pytest.raises(G4SError):
        plan.spmm(B.data_ptr(), B.data_ptr())
