"""CPU model of the render kernels' cooperative cull (csrc/render_common.cuh: block_mask8): the 8-bit mask computed per
(tile, Gaussian) must be CONSERVATIVE -- every pixel of the tile at which the Gaussian passes the alpha >= 1/255 test
(q(d) <= tau) lies in a warp pixel block whose bit is set -- and tight (a bit is set only if the continuous ellipse
reaches the block's rectangle).  The formula is restated in numpy, operation for operation."""
import numpy as np


def block_mask8(gx, gy, tau, A, B, C, tx0, ty0):
    """numpy restatement of ggrt::block_mask8 (float32 arithmetic, same order; rcp_approx -> exact reciprocal)."""
    f = np.float32
    gx, gy, tau, A, B, C, tx0, ty0 = map(f, (gx, gy, tau, A, B, C, tx0, ty0))
    iA, iC, B2 = f(1) / A, f(1) / C, f(2) * B
    lox, hix, dxe, qxe, dy0, bxe = [], [], [], [], [], []
    for c in range(2):
        lo = tx0 + f(8 * c) - gx
        hi = lo + f(7)
        d = min(max(f(0), lo), hi)
        lox.append(lo), hix.append(hi), dxe.append(d)
        qxe.append(A * d * d), bxe.append(B2 * d), dy0.append(-B * d * iC)
    mask = 0
    for r in range(4):
        loy = ty0 + f(4 * r) - gy
        hiy = loy + f(3)
        dye = min(max(f(0), loy), hiy)
        qye, bye, dx0 = C * dye * dye, B2 * dye, -B * dye * iA
        for c in range(2):
            dy = min(max(dy0[c], loy), hiy)
            dx = min(max(dx0, lox[c]), hix[c])
            qv = (C * dy + bxe[c]) * dy + qxe[c]
            qh = (A * dx + bye) * dx + qye
            if min(qv, qh) <= tau:
                mask |= 1 << (2 * r + c)
    return mask


def test_block_mask_is_conservative_and_tight():
    rng = np.random.default_rng(0)
    tx0, ty0 = 32.0, 48.0
    px, py = np.meshgrid(np.arange(16) + tx0, np.arange(16) + ty0)
    blk = (((py - ty0) // 4) * 2 + (px - tx0) // 8).astype(int)  # block id of every pixel of the tile
    n_checked = n_bits = n_tight = 0
    for _ in range(4000):
        # a random positive-definite conic (inverse 2D covariance with the +0.3 low pass), centre around the tile
        s1, s2 = rng.uniform(0.3, 40.0, 2)
        th = rng.uniform(0, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        cov = R @ np.diag([s1, s2]) @ R.T + 0.3 * np.eye(2)
        con = np.linalg.inv(cov)
        A, B, C = np.float32(con[0, 0]), np.float32(con[0, 1]), np.float32(con[1, 1])
        gx, gy = rng.uniform(tx0 - 12, tx0 + 28), rng.uniform(ty0 - 12, ty0 + 28)
        o = rng.uniform(0.01, 1.0)
        tau = 2.0 * np.log(255.0 * o)
        if tau <= 0:
            continue
        tau_c = np.float32(tau * 1.001 + 0.02)  # the inflated threshold geometry_kernel stores
        m = block_mask8(gx, gy, tau_c, A, B, C, tx0, ty0)
        dx, dy = np.float32(gx) - px.astype(np.float32), np.float32(gy) - py.astype(np.float32)
        q = A * dx * dx + 2 * B * dx * dy + C * dy * dy  # alpha >= 1/255  <=>  q <= tau
        hit_blocks = set(blk[q <= tau].tolist())
        for b in hit_blocks:  # conservative: no contributing pixel is ever culled
            assert m >> b & 1, (gx, gy, A, B, C, tau, b, m)
        # tight: a set bit means the continuous ellipse (inflated threshold) reaches the block's rectangle
        for b in range(8):
            if m >> b & 1:
                n_bits += 1
                x0, y0 = tx0 + 8 * (b & 1), ty0 + 4 * (b >> 1)
                xs, ys = np.meshgrid(np.linspace(x0, x0 + 7, 57), np.linspace(y0, y0 + 3, 25))
                ddx, ddy = gx - xs, gy - ys
                qq = con[0, 0] * ddx * ddx + 2 * con[0, 1] * ddx * ddy + con[1, 1] * ddy * ddy
                n_tight += bool((qq <= tau_c * 1.02 + 0.05).any())
        n_checked += 1
    assert n_checked > 3000 and n_bits > 3000
    assert n_tight >= 0.999 * n_bits  # (dense sampling of the rectangle finds the ellipse wherever a bit is set)
