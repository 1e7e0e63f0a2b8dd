"""CPU restatement of the ADDRESSING of the materialised attention backward (commu-code_b200/csrc/attn_bwd_mat.cu):
the coarse-sheared dS workspace, the residue-class view, the reversed distance table and the tile walks of the three
band GEMMs, emulated with numpy exactly as the kernels issue their TMA boxes (out-of-bounds = zero fill).  The sums
must equal the plain definitions  dq_A = dS K,  dq_C[i] = sum_j dS[i,j] R[i+M-j],  dR[delta] = sum_{i} dS[i, i+M-delta] qv[i].
No arithmetic of the product is exercised here - this pins the index conventions (it caught nothing is not the point:
the kernels are written against these formulas)."""
import numpy as np
import pytest


def geometry(T, M):
    Tpad = (T + 127) // 128 * 128
    Kp = (T + M + 127) // 128 * 128
    X = Tpad
    P = Kp + Tpad + 64
    Y0 = M + X + 8
    return Tpad, Kp, X, P, Y0


def box2d(buf, col, row, w, h):
    """TMA tiled load of a {w cols, h rows} box at (col, row) with zero fill outside the tensor."""
    out = np.zeros((h, w), buf.dtype)
    R, C = buf.shape
    for rr in range(h):
        r = row + rr
        if r < 0 or r >= R:
            continue
        c0, c1 = max(col, 0), min(col + w, C)
        if c1 > c0:
            out[rr, c0 - col:c1 - col] = buf[r, c0:c1]
    return out


@pytest.mark.parametrize("T,M", [(300, 200), (128, 0), (64, 77), (384, 384)])
def test_coarse_sheared_layout_and_band_gemms(T, M):
    rng = np.random.default_rng(T + M)
    K = T + M
    Tpad, Kp, X, P, Y0 = geometry(T, M)
    D = 4                                         # head dim stand-in
    ii, jj = np.arange(T)[:, None], np.arange(K)[None, :]
    dS = rng.standard_normal((T, K)) * (jj <= ii + M)          # causal support
    Kmat = rng.standard_normal((K, D))
    R = rng.standard_normal((K, D))                              # by distance, kr = K
    qv = rng.standard_normal((T, D))
    # ---- pass 1: every causal 128 x 128 tile is stored as 16 boxes {64 cols, 8 rows} per key half ----
    ws = np.zeros((Tpad, P))
    dS_pad = np.zeros((Tpad, Kp))
    dS_pad[:T, :K] = dS
    for jt in range(Kp // 128):
        j0 = jt * 128
        if j0 >= K:
            continue
        for it in range(max(0, j0 - M) // 128, (T - 1) // 128 + 1):
            i0 = it * 128
            for half in range(2):
                for g in range(16):
                    col, row = j0 + 64 * half + X - (i0 + 8 * g), i0 + 8 * g
                    assert 8 <= col and col + 64 <= P - 8
                    ws[row:row + 8, col:col + 64] = dS_pad[row:row + 8, j0 + 64 * half:j0 + 64 * half + 64]
    # ---- mode A: dq_A rows i0..i0+127 ----
    dqa = np.zeros((Tpad, D))
    for it in range((T + 127) // 128):
        i0 = it * 128
        nsteps = (min(T - 1, i0 + 127) + M) // 128 + 1
        for s in range(nsteps):
            j0 = s * 128
            tile = np.zeros((128, 128))
            for half in range(2):
                for g in range(16):
                    tile[8 * g:8 * g + 8, 64 * half:64 * half + 64] = box2d(ws, j0 + 64 * half + X - (i0 + 8 * g), i0 + 8 * g, 64, 8)
            kt = box2d(Kmat, 0, j0, D, 128)
            dqa[i0:i0 + 128] += tile @ kt
    assert np.allclose(dqa[:T], dS @ Kmat)
    # ---- reversed table ----
    rrev = np.zeros((Y0 + 1, D))
    for y in range(Y0 + 1):
        d = Y0 - y
        if 0 <= d < K:
            rrev[y] = R[d]
    # residue view: row (a, r) = workspace row 8a + r, same columns
    nab = (Tpad // 8 + 127) // 128
    amax = Tpad // 8 - 1

    def a_tile(c, r, a0):
        out = np.zeros((128, 128))
        for a in range(128):
            if a0 + a <= amax:
                out[a] = box2d(ws, c, 8 * (a0 + a) + r, 128, 1)[0]
        return out
    # ---- mode C ----
    dqc = np.zeros((Tpad, D))
    for ab in range(nab):
        a0 = ab * 128
        for r in range(8):
            lo = X - 8 * min(a0 + 127, amax)
            s_first = max(lo, 0) // 128
            nsteps = (X + M + 7) // 128 - s_first + 1
            acc = np.zeros((128, D))
            for s in range(nsteps):
                cc = (s_first + s) * 128
                acc += a_tile(cc, r, a0) @ box2d(rrev, 0, cc + 8 - r, D, 128)
            for li in range(128):
                i = 8 * (a0 + li) + r
                if i < T:
                    dqc[i] = acc[li]
    ref_c = np.zeros((T, D))
    for i in range(T):
        for j in range(min(K, i + M + 1)):
            ref_c[i] += dS[i, j] * R[i + M - j]
    assert np.allclose(dqc[:T], ref_c)
    # ---- mode R ----
    dR = np.zeros((K, D))
    ncb = (X + M + 7) // 128 + 1
    for cb in range(ncb):
        c0 = cb * 128
        for r in range(8):
            t = X - c0 - 127
            a_first = (t + 7) // 8 if t > 0 else 0         # the sweep starts at the first ROW that holds data here;
            rows = Tpad // 8                                # rows past the end are TMA zero fill (a_tile / ct below)
            nper = (rows - a_first + 127) // 128 if a_first < rows else 0
            acc = np.zeros((128, D))
            for st in range(nper):
                a0 = a_first + st * 128
                ct = np.zeros((128, D))                     # (q+v) rows 8(a0+k)+r, zero beyond T
                for k in range(128):
                    i = 8 * (a0 + k) + r
                    if i < T:
                        ct[k] = qv[i]
                acc += a_tile(c0, r, a0).T @ ct
            for li in range(128):
                delta = r + M + X - (c0 + li)
                if 0 <= delta < K:
                    dR[delta] += acc[li]
    ref_r = np.zeros((K, D))
    for i in range(T):
        for j in range(min(K, i + M + 1)):
            ref_r[i + M - j] += dS[i, j] * qv[i]
    assert np.allclose(dR, ref_r)


def merged_ticket_tile(ticket, Tpad):
    """relattn_bwd_band_ac_kernel: ticket -> ((b,h) index, role, tile, hand-over block, number of dq_A tiles of the block)."""
    nab = (Tpad // 8 + 127) // 128
    nta = Tpad // 128
    per_bh = nta + 8 * nab
    bh, u = divmod(ticket, per_bh)
    ab = nab - 1
    while ab >= 0:
        na = min(8, nta - 8 * ab)
        if u < na:
            return bh, "A", 8 * ab + na - 1 - u, ab, na
        u -= na
        if u < 8:
            return bh, "C", (ab, u), ab, na
        u -= 8
        ab -= 1
    raise AssertionError("ticket outside the grid")


@pytest.mark.parametrize("T", [1, 100, 128, 300, 1024, 1100, 2048, 4096])
def test_merged_dq_launch_ticket_order(T):
    """The merged dq_A / dq_C launch: every row tile and every (row block, residue) is issued exactly once per (b,h); the
    dq_A tiles a dq_C tile waits for cover exactly its rows and hold LOWER tickets (they run or are done when it waits)."""
    Tpad = (T + 127) // 128 * 128
    nab = (Tpad // 8 + 127) // 128
    nta = Tpad // 128
    per_bh = nta + 8 * nab
    BH = 3
    seen_a, seen_c, first_c_ticket, a_tickets = set(), set(), {}, {}
    for ticket in range(BH * per_bh):
        bh, role, tile, ab, na = merged_ticket_tile(ticket, Tpad)
        assert bh == ticket // per_bh
        if role == "A":
            assert 0 <= tile < nta and (bh, tile) not in seen_a
            assert tile // 8 == ab                      # the row tile signals the flag of ITS 1024-row block
            seen_a.add((bh, tile))
            a_tickets.setdefault((bh, ab), []).append(ticket)
        else:
            assert (bh, tile) not in seen_c
            seen_c.add((bh, tile))
            first_c_ticket.setdefault((bh, ab), ticket)
            assert na == len(a_tickets[(bh, ab)])       # the count it waits for == dq_A tiles of the block, all issued earlier
            assert max(a_tickets[(bh, ab)]) < ticket
            # its rows 8 (128 ab + l) + r, l < 128, lie in the row tiles 8 ab .. 8 ab + 7 (those that exist)
            lo, hi = 1024 * ab, min(1024 * ab + 1023, Tpad - 1)
            assert {(bh, t) for t in range(lo // 128, hi // 128 + 1)} <= seen_a
    assert len(seen_a) == BH * nta and len(seen_c) == BH * nab * 8
