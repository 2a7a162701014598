"""CPU check of the SH term list the CUDA kernels share (csrc/common.cuh: GGRT_SH_TERMS_0..4): every basis function
B_k against the oracle's basis, and every derivative (dB_k/dx, dB_k/dy, dB_k/dz) against autograd of the oracle's
basis -- the colour kernel contracts these derivatives with the SH coefficients into the Jacobian the backward uses
instead of re-reading the SH table, so a typo in one of the 75 expressions would be a silent gradient error."""
import re
from pathlib import Path

import torch

from oracle.torch_ref import sh_basis

ROOT = Path(__file__).resolve().parent.parent
SRC = (ROOT / "ggrt_official_b200" / "csrc" / "common.cuh").read_text()


def _constants():
    return {m.group(1): float(m.group(2)) for m in re.finditer(r"#define (GGRT_SH_C\w+) (-?[0-9.]+)f", SRC)}


def _terms():
    """[(k, B, BX, BY, BZ)] as Python expression strings, parsed from the T(...) entries of the macro bodies."""
    body = "".join(re.findall(r"#define GGRT_SH_TERMS_\d\(T\)(.*?)(?=\n#define|\n//)", SRC, flags=re.S))
    body = body.replace("\\\n", " ")
    out, i = [], 0
    while True:
        i = body.find("T(", i)
        if i < 0:
            break
        depth, j = 0, i + 1
        while True:  # matching parenthesis of this T( ... )
            depth += body[j] == "("
            depth -= body[j] == ")"
            if depth == 0:
                break
            j += 1
        args, cur, depth = [], "", 0
        for ch in body[i + 2: j]:
            if ch == "," and depth == 0:
                args.append(cur.strip())
                cur = ""
            else:
                depth += ch == "("
                depth -= ch == ")"
                cur += ch
        args.append(cur.strip())
        assert len(args) == 5, args
        out.append(tuple(args))
        i = j
    return out


def _py(expr: str) -> str:
    return re.sub(r"(\d+\.\d*)f", r"\1", expr).replace("GGRT_Z", "0.0")


def test_term_list_matches_the_oracle_basis_and_its_derivatives():
    terms = _terms()
    assert [int(t[0]) for t in terms] == list(range(25))
    torch.manual_seed(0)
    d = torch.randn(64, 3, dtype=torch.float64)
    d = (d / d.norm(dim=1, keepdim=True)).requires_grad_()
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    env = dict(_constants(), x=x, y=y, z=z, xx=x * x, yy=y * y, zz=z * z, xy=x * y, yz=y * z, xz=x * z)
    ref = sh_basis(4, d)  # [64, 25]
    for k, B, BX, BY, BZ in terms:
        k = int(k)
        val = eval(_py(B), {}, env)
        val = val if torch.is_tensor(val) else torch.full_like(x, float(val))
        assert torch.allclose(val, ref[:, k], atol=1e-12), (k, B)
        # derivative of the POLYNOMIAL w.r.t. the free variables (x, y, z), as the kernels use it (the normalisation of
        # the direction is differentiated separately)
        (g,) = torch.autograd.grad(ref[:, k].sum(), d, retain_graph=True)
        for axis, e in enumerate((BX, BY, BZ)):
            dv = eval(_py(e), {}, env)
            dv = dv if torch.is_tensor(dv) else torch.full_like(x, float(dv))
            assert torch.allclose(dv, g[:, axis], atol=1e-10), (k, "xyz"[axis], e)


def test_direction_is_invariant_under_the_per_view_scene_scale():
    """graph.CapturedViews merges views whose scene scale s_v lives on the device (scale-invariant rendering: s_v = 1 /
    near_v) with ONE host-side scale of 1: basis(normalize(s m - c)) == basis(normalize(m - c / s)) for s > 0, so the
    camera centres are divided by their view's scale instead."""
    torch.manual_seed(3)
    m = torch.randn(200, 3, dtype=torch.float64) * 3
    acc_a = acc_b = 0
    for v in range(4):
        s = 0.3 + 1.7 * torch.rand((), dtype=torch.float64)
        c = torch.randn(3, dtype=torch.float64)
        d = torch.randn(200, 3, dtype=torch.float64)
        da = s * m - c
        db = m - c / s
        ba = sh_basis(4, da / da.norm(dim=1, keepdim=True))
        bb = sh_basis(4, db / db.norm(dim=1, keepdim=True))
        acc_a = acc_a + ba[:, :, None] * d[:, None, :]
        acc_b = acc_b + bb[:, :, None] * d[:, None, :]
    assert torch.allclose(acc_a, acc_b, atol=1e-12)
