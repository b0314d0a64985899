"""CPU emulation of the ldmatrix / mma.m16n8k16 fragment algebra used by csrc/temporal_attn_mma.cu.

The experimental tensor-core temporal attention could not be run on hardware in round 1, so its index algebra —
which lane addresses which shared-memory row, how C fragments are re-packed as A fragments, which products need the
.trans form — is checked here against numpy with a lane-by-lane emulator written from the PTX ISA description of
ldmatrix (m8n8 .x2/.x4, .trans) and mma.sync.aligned.m16n8k16.row.col. The functions below mirror the device code
line by line (same address formulas); they are test infrastructure only.
"""
import numpy as np
import pytest

T16, D = 16, 64


def ldmatrix(tile, addrs, n, trans):
    """tile: 2-D array; addrs[lane] = (row, col) start of an 8-element row; matrices i = 0..n-1 use lanes 8i..8i+7.
    Returns regs[lane][i] = (e0, e1)."""
    regs = [[None] * n for _ in range(32)]
    for i in range(n):
        m = np.stack([tile[addrs[8 * i + r][0], addrs[8 * i + r][1]:addrs[8 * i + r][1] + 8] for r in range(8)])
        for lane in range(32):
            g, q = lane // 4, lane % 4
            regs[lane][i] = (m[2 * q, g], m[2 * q + 1, g]) if trans else (m[g, 2 * q], m[g, 2 * q + 1])
    return regs


def mma16816(c, a, b):
    """c[lane] = 4 floats, a[lane] = 4 regs of 2, b[lane] = 2 regs of 2  ->  c += A(16x16) @ B(16x8)."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8)); C = np.zeros((16, 8))
    for lane in range(32):
        g, q = lane // 4, lane % 4
        for e in range(2):
            A[g, 2 * q + e] = a[lane][0][e]; A[g + 8, 2 * q + e] = a[lane][1][e]
            A[g, 8 + 2 * q + e] = a[lane][2][e]; A[g + 8, 8 + 2 * q + e] = a[lane][3][e]
            B[2 * q + e, g] = b[lane][0][e]; B[8 + 2 * q + e, g] = b[lane][1][e]
            C[g, 2 * q + e] = c[lane][e]; C[g + 8, 2 * q + e] = c[lane][2 + e]
    Dm = A @ B + C
    for lane in range(32):
        g, q = lane // 4, lane % 4
        c[lane] = [Dm[g, 2 * q], Dm[g, 2 * q + 1], Dm[g + 8, 2 * q], Dm[g + 8, 2 * q + 1]]


# ---- the device helpers, address formula for address formula ----
def load_a(m, col0):
    return ldmatrix(m, [((l & 15), col0 + (l >> 4) * 8) for l in range(32)], 4, False)


def load_b_rows(m, n0, k0):
    return ldmatrix(m, [(n0 + (l & 7), k0 + ((l >> 3) & 1) * 8) for l in range(32)], 2, False)


def load_b_cols(m, n0):
    return ldmatrix(m, [((l & 15), n0) for l in range(32)], 2, True)


def zeros_c():
    return [[0.0] * 4 for _ in range(32)]


def pack_a(c0, c1):
    """C fragments of n-tiles 0 and 1 -> A fragment (device: pack_a)."""
    return [[(c0[l][0], c0[l][1]), (c0[l][2], c0[l][3]), (c1[l][0], c1[l][1]), (c1[l][2], c1[l][3])] for l in range(32)]


def frag16x16(c0, c1):
    M = np.zeros((16, 16))
    for l in range(32):
        g, q = l // 4, l % 4
        for nt, c in enumerate((c0, c1)):
            M[g, nt * 8 + 2 * q] = c[l][0]; M[g, nt * 8 + 2 * q + 1] = c[l][1]
            M[g + 8, nt * 8 + 2 * q] = c[l][2]; M[g + 8, nt * 8 + 2 * q + 1] = c[l][3]
    return M


def frag16x64(cs):
    M = np.zeros((16, 64))
    for l in range(32):
        g, q = l // 4, l % 4
        for nt in range(8):
            M[g, nt * 8 + 2 * q] = cs[nt][l][0]; M[g, nt * 8 + 2 * q + 1] = cs[nt][l][1]
            M[g + 8, nt * 8 + 2 * q] = cs[nt][l][2]; M[g + 8, nt * 8 + 2 * q + 1] = cs[nt][l][3]
    return M


def scores(Q, K):
    s = [zeros_c(), zeros_c()]
    for kk in range(4):
        a = load_a(Q, kk * 16)
        for nt in range(2):
            mma16816(s[nt], a, load_b_rows(K, nt * 8, kk * 16))
    return s


def to_frags(M):
    """16x16 matrix -> (c0, c1) C fragments."""
    cs = [zeros_c(), zeros_c()]
    for l in range(32):
        g, q = l // 4, l % 4
        for nt in range(2):
            cs[nt][l] = [M[g, nt * 8 + 2 * q], M[g, nt * 8 + 2 * q + 1], M[g + 8, nt * 8 + 2 * q], M[g + 8, nt * 8 + 2 * q + 1]]
    return cs


@pytest.mark.parametrize('T', [16, 11])
def test_fragment_algebra_matches_numpy(T):
    rng = np.random.default_rng(0)
    Q, K, V, dO = (np.zeros((16, D)) for _ in range(4))
    for M in (Q, K, V, dO):
        M[:T] = rng.standard_normal((T, D))
    scale = 0.37
    # ---- forward: S = Q K^T, causal softmax, O = P V ----
    s = scores(Q, K)
    S = frag16x16(*s)
    np.testing.assert_allclose(S, Q @ K.T, atol=1e-9)
    mask = (np.arange(16)[None, :] <= np.arange(16)[:, None]) & (np.arange(16)[None, :] < T)
    Sm = np.where(mask, S * scale, -np.inf)
    m = Sm.max(1, keepdims=True); m[np.isinf(m)] = 0
    Pu = np.exp(Sm - m)
    l = Pu.sum(1, keepdims=True)
    P = np.where(l > 0, Pu / np.where(l > 0, l, 1), 0)
    pa = pack_a(*to_frags(Pu))                                   # un-normalised P as the A operand
    o = []
    for nt in range(8):
        c = zeros_c()
        mma16816(c, pa, load_b_cols(V, nt * 8))
        o.append(c)
    O = frag16x64(o) * np.where(l > 0, 1 / np.where(l > 0, l, 1), 0)
    ref_O = P @ V
    np.testing.assert_allclose(O[:T], ref_O[:T], atol=1e-9)
    # ---- backward ----
    dp = [zeros_c(), zeros_c()]
    for kk in range(4):
        a = load_a(dO, kk * 16)
        for nt in range(2):
            mma16816(dp[nt], a, load_b_rows(V, nt * 8, kk * 16))
    dP = frag16x16(*dp)
    np.testing.assert_allclose(dP, dO @ V.T, atol=1e-9)
    delta = (P * dP).sum(1, keepdims=True)
    dS = P * (dP - delta) * scale
    dsa = pack_a(*to_frags(dS))
    dq = []
    for nt in range(8):
        c = zeros_c()
        mma16816(c, dsa, load_b_cols(K, nt * 8))
        dq.append(c)
    np.testing.assert_allclose(frag16x64(dq), dS @ K, atol=1e-9)
    # transposed A operands from the [query][key] tiles (device: ldsm_x4_trans with r, c as below)
    addrs = [((lane & 7) + ((lane >> 4) & 1) * 8, ((lane >> 3) & 1) * 8) for lane in range(32)]
    pta = ldmatrix(P, addrs, 4, True)
    dsta = ldmatrix(dS, addrs, 4, True)
    dv, dk = [], []
    for nt in range(8):
        c = zeros_c(); mma16816(c, pta, load_b_cols(dO, nt * 8)); dv.append(c)
        c = zeros_c(); mma16816(c, dsta, load_b_cols(Q, nt * 8)); dk.append(c)
    np.testing.assert_allclose(frag16x64(dv), P.T @ dO, atol=1e-9)
    np.testing.assert_allclose(frag16x64(dk), dS.T @ Q, atol=1e-9)
    # and the products are the attention gradients (numpy autograd-free check against finite differences of O)
    eps = 1e-6
    Qp = Q.copy(); Qp[2, 5] += eps
    Sp = np.where(mask, (Qp @ K.T) * scale, -np.inf)
    mp = Sp.max(1, keepdims=True); mp[np.isinf(mp)] = 0
    Pp = np.exp(Sp - mp); lp = Pp.sum(1, keepdims=True); Pp = np.where(lp > 0, Pp / np.where(lp > 0, lp, 1), 0)
    fd = (((Pp @ V) - ref_O) * dO).sum() / eps
    assert abs(fd - (dS @ K)[2, 5]) < 1e-4 * max(1.0, abs(fd))
