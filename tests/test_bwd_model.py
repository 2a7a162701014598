"""CPU model of the render-backward kernel's two reformulations (csrc/render_bwd.cu), checked against the plain
back-to-front loop of A.4:
  * the alpha-blend replay as an associative scan of the maps (T, -R) -> (a T, -R + nb T), with a lane composing its two
    Gaussians, a Kogge-Stone scan over the 8 pairs and the back Gaussian's prefix recovered by undoing the front one;
  * the gradient sums as [Gaussians x pixels] . [pixels x monomials] products about the block centre, shifted to the
    Gaussian's own centre afterwards, with tf32 head / tail splitting of the left operand."""
import numpy as np


def _tf32_head(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def test_pair_scan_reproduces_the_sequential_replay():
    rng = np.random.default_rng(1)
    for _ in range(200):
        alpha = rng.uniform(0, 0.99, 16) * (rng.uniform(size=16) < 0.7)  # slot 0 = backmost; inactive ones have alpha 0
        sdot = rng.normal(size=16)                                       # c_i . g at this pixel
        Tb, nRb = rng.uniform(1e-3, 1.0), rng.normal()                   # state behind the chunk: T, -(sum behind)
        # plain loop, back to front: T_i = T_behind / (1 - alpha_i); -(sum incl. i) = -(sum behind) - alpha_i T_i sdot_i
        T, nR, T_ref, nR_ref = Tb, nRb, [], []
        for i in range(16):
            T = T / (1.0 - alpha[i])
            nR = nR - alpha[i] * T * sdot[i]
            T_ref.append(T), nR_ref.append(nR)
        # kernel formulation
        a = 1.0 / (1.0 - alpha)
        nb = -alpha * sdot * a
        e, o = slice(0, 16, 2), slice(1, 16, 2)                          # a lane's back / front Gaussian
        A = a[e] * a[o]
        nB = nb[o] * a[e] + nb[e]                                        # (a_o a_e, nb_o a_e + nb_e)
        for d in (1, 2, 4):                                              # inclusive scan over the 8 pairs
            Ap, Bp = np.ones(8), np.zeros(8)
            Ap[d:], Bp[d:] = A[:-d], nB[:-d]                             # identity where no lane d below exists
            A, nB = A * Ap, nB * Ap + Bp
        Ae = A * (1.0 - alpha[o])                                        # undo the front Gaussian: 1 / a_o = 1 - alpha_o
        nBe = nB - nb[o] * Ae
        T_k = np.empty(16)
        nR_k = np.empty(16)
        T_k[o], nR_k[o] = Tb * A, Tb * nB + nRb
        T_k[e], nR_k[e] = Tb * Ae, Tb * nBe + nRb
        assert np.allclose(T_k, T_ref, rtol=1e-10) and np.allclose(nR_k, nR_ref, rtol=1e-9, atol=1e-12)


def test_block_centred_moments_shift_to_the_gaussian_centre():
    rng = np.random.default_rng(2)
    bx0, by0 = 40.0, 72.0
    px, py = np.meshgrid(bx0 + np.arange(8), by0 + np.arange(4))
    x, y = px - (bx0 + 3.5), py - (by0 + 1.5)                            # block-centred pixel coordinates
    for _ in range(200):
        q = rng.normal(size=(4, 8)) * (rng.uniform(size=(4, 8)) < 0.5)   # G dL/dalpha per pixel (sparse)
        gx, gy = rng.uniform(bx0 - 20, bx0 + 28), rng.uniform(by0 - 20, by0 + 24)
        dx, dy = gx - px, gy - py
        ref = [q.sum(), (q * dx).sum(), (q * dy).sum(), (q * dx * dx).sum(), (q * dx * dy).sum(), (q * dy * dy).sum()]
        S1, Sx, Sy, Sxx, Sxy, Syy = [(q * m).sum() for m in (np.ones_like(x), x, y, x * x, x * y, y * y)]
        u, w = gx - (bx0 + 3.5), gy - (by0 + 1.5)
        mx, my = u * S1 - Sx, w * S1 - Sy
        got = [S1, mx, my, u * mx - u * Sx + Sxx, w * mx - u * Sy + Sxy, w * my - w * Sy + Syy]
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_tf32_head_tail_split_keeps_22_bits():
    rng = np.random.default_rng(3)
    q = (rng.normal(size=100000) * 10.0 ** rng.uniform(-6, 3, 100000)).astype(np.float32)
    head = _tf32_head(q)
    tail = q - head                       # exact in float32
    assert np.array_equal(head + tail, q)
    used = head.astype(np.float64) + _tf32_head(tail).astype(np.float64)  # what two tf32 MMAs see (truncated inputs)
    rel = np.abs(used - q.astype(np.float64)) / np.abs(q.astype(np.float64))
    assert rel.max() <= 2.0 ** -20
    # the coordinate monomials of the block (+-3.5, +-1.5 and their products) are exact in tf32
    xs, ys = np.arange(8) - 3.5, np.arange(4) - 1.5
    mono = np.float32(np.concatenate([xs, ys, np.outer(xs, xs).ravel(), np.outer(xs, ys).ravel(), np.outer(ys, ys).ravel()]))
    assert np.array_equal(_tf32_head(mono), mono)
