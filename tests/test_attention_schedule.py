"""CPU models of two pieces of device-side index logic whose mistakes would be silent on most shapes:

* the work-item schedule of the tcgen05 attention kernel (csrc/k_attn_tc.cu: launch_attn_tc's n_full / n_items
  split, get_item, ItemWalk): every (clip, 128-row query tile) must be visited exactly once, by exactly one
  CTA, for any batch / length / grid -- including the single-tile tail items and the staggered order;
* the fragment mapping of the T <= 8 kernel (csrc/k_attn_small.cu): the permuted contraction / output
  dimension orders and the block-diagonal P operand must reproduce softmax(QK^T/sqrt(d))V.
These are restatements of the kernels' arithmetic in NumPy (no GPU); the GPU parity tests check the kernels."""
import numpy as np
import pytest
import torch

BM, BKV = 128, 64


def _schedule(B, T, num_sms=148, stagger=1):
    npairs = (T + 2 * BM - 1) // (2 * BM)
    n_pairs = B * npairs
    grid = min(n_pairs, num_sms)
    n_full = n_items = n_pairs
    rem = n_pairs % grid
    if n_pairs > grid and rem > 0 and 2 * rem <= grid:
        n_full, n_items = n_pairs - rem, n_pairs - rem + 2 * rem

    def get_item(idx):
        pair, tile = idx, -1
        if idx >= n_full:
            pair, tile = n_full + ((idx - n_full) >> 1), (idx - n_full) & 1
        b, q0 = pair // npairs, (pair % npairs) * 2 * BM
        ntile = 2 if q0 + BM < T else 1
        valid = True
        if tile >= 0:
            if tile == 1 and ntile == 1:
                valid = False
            q0 += tile * BM
            ntile = 1
        return b, q0, ntile, valid

    visits = {}
    per_cta = []
    for cta in range(grid):
        n = (n_items - cta + grid - 1) // grid if cta < n_items else 0
        last = cta + (n - 1) * grid
        rot = 1 if (stagger and n > 1 and (cta & 1) and last >= n_full) else 0
        order = []
        for r in range(n):
            rr = (n - 1 if r == 0 else r - 1) if rot else r
            order.append(cta + rr * grid)
        assert sorted(order) == list(range(cta, n_items, grid))          # a permutation of the plain walk
        per_cta.append(order)
        for idx in order:
            b, q0, ntile, valid = get_item(idx)
            if not valid:
                continue
            for t in range(ntile):
                key = (b, q0 + t * BM)
                visits[key] = visits.get(key, 0) + 1
    return visits, per_cta, n_full


@pytest.mark.parametrize("B,T", [(1, 1), (1, 64), (1, 128), (1, 129), (3, 300), (5, 256), (2, 512), (7, 513),
                                 (256, 512), (255, 512), (300, 512), (64, 8192), (149, 256), (148, 257),
                                 (1000, 7), (37, 1000)])
@pytest.mark.parametrize("stagger", [0, 1])
def test_every_query_tile_visited_exactly_once(B, T, stagger):
    visits, _, _ = _schedule(B, T, stagger=stagger)
    want = {(b, q0) for b in range(B) for q0 in range(0, T, BM)}
    assert set(visits) == want
    assert all(v == 1 for v in visits.values())


def test_staggered_order_puts_the_tail_item_first_on_odd_ctas():
    _, per_cta, n_full = _schedule(256, 512, stagger=1)        # BASELINE config 2: 444 pairs + 136 single tiles
    assert n_full == 444
    for cta, order in enumerate(per_cta):
        has_tail = order and max(order) >= n_full
        if has_tail and (cta & 1):
            assert order[0] >= n_full and order[1:] == sorted(order[1:])
        else:
            assert order == sorted(order)


# ---------------------------------------------------------------------------------------------
def _mma_16816(d, A, Bm):
    """mma.sync.m16n8k16 row.col with the PTX fragment layout: lane = 4 g + t4;
    A regs: (g, 2t4..), (g+8, 2t4..), (g, 2t4+8..), (g+8, 2t4+8..); B regs: (k=2t4.., n=g), (k=2t4+8.., n=g);
    C regs: (g, 2t4), (g, 2t4+1), (g+8, 2t4), (g+8, 2t4+1)."""
    Am, Bk = np.zeros((16, 16)), np.zeros((16, 8))
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        Am[g, 2 * t4:2 * t4 + 2], Am[g + 8, 2 * t4:2 * t4 + 2] = A[lane][0], A[lane][1]
        Am[g, 2 * t4 + 8:2 * t4 + 10], Am[g + 8, 2 * t4 + 8:2 * t4 + 10] = A[lane][2], A[lane][3]
        Bk[2 * t4:2 * t4 + 2, g], Bk[2 * t4 + 8:2 * t4 + 10, g] = Bm[lane][0], Bm[lane][1]
    C = Am @ Bk
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        d[lane] += [C[g, 2 * t4], C[g, 2 * t4 + 1], C[g + 8, 2 * t4], C[g + 8, 2 * t4 + 1]]


def _attn_small_model(Q, K, V, lengths, T):
    n_win, D = Q.shape[0], 128
    Qf, Kf, Vf = Q.reshape(-1, D), K.reshape(-1, D), V.reshape(-1, D)
    O = np.full((n_win * T, D), np.nan)
    c = 1.4426950408889634 * 0.08838834764831845
    for pair in range((n_win + 1) // 2):
        wa = 2 * pair
        has_b = wa + 1 < n_win
        wb = wa + 1 if has_b else wa
        ra, rb = wa * T, wb * T
        la, lb = min(max(int(lengths[wa]), 0), T), min(max(int(lengths[wb]), 0), T)
        sa, sb = np.zeros((32, 4)), np.zeros((32, 4))
        for s in range(8):
            A, Ba, Bb = [], [], []
            for lane in range(32):
                g, t4 = lane >> 2, lane & 3
                gq = min(g, T - 1)
                qa = Qf[ra + gq, t4 * 32:(t4 + 1) * 32].reshape(16, 2)
                qb = Qf[rb + gq, t4 * 32:(t4 + 1) * 32].reshape(16, 2)
                ka = Kf[ra + gq, t4 * 32:(t4 + 1) * 32].reshape(16, 2)
                kb = Kf[rb + gq, t4 * 32:(t4 + 1) * 32].reshape(16, 2)
                A.append([qa[2 * s], qb[2 * s], qa[2 * s + 1], qb[2 * s + 1]])
                Ba.append([ka[2 * s], ka[2 * s + 1]])
                Bb.append([kb[2 * s], kb[2 * s + 1]])
            _mma_16816(sa, A, Ba)
            _mma_16816(sb, A, Bb)
        P = np.zeros((32, 4))
        for quad in range(8):
            xs, ys = [], []
            for t4 in range(4):
                lane = quad * 4 + t4
                xs += [sa[lane][0] * c if 2 * t4 < la else -np.inf, sa[lane][1] * c if 2 * t4 + 1 < la else -np.inf]
                ys += [sb[lane][2] * c if 2 * t4 < lb else -np.inf, sb[lane][3] * c if 2 * t4 + 1 < lb else -np.inf]
            xs, ys = np.array(xs), np.array(ys)
            pa, pb = np.exp2(xs - xs.max()), np.exp2(ys - ys.max())
            pa, pb = pa / pa.sum(), pb / pb.sum()
            for t4 in range(4):
                P[quad * 4 + t4] = [pa[2 * t4], pa[2 * t4 + 1], pb[2 * t4], pb[2 * t4 + 1]]
        for m in range(16):
            A, Bm = [], []
            for lane in range(32):
                g, t4 = lane >> 2, lane & 3
                k0, k1 = min(2 * t4, T - 1), min(2 * t4 + 1, T - 1)
                A.append([P[lane][0:2], np.zeros(2), np.zeros(2), P[lane][2:4]])
                Bm.append([np.array([Vf[ra + k0, g * 16 + m], Vf[ra + k1, g * 16 + m]]),
                           np.array([Vf[rb + k0, g * 16 + m], Vf[rb + k1, g * 16 + m]])])
            d = np.zeros((32, 4))
            _mma_16816(d, A, Bm)
            for lane in range(32):
                g, t4 = lane >> 2, lane & 3
                if g < T:
                    O[ra + g, t4 * 32 + m], O[ra + g, t4 * 32 + 16 + m] = d[lane][0], d[lane][1]
                    if has_b:
                        O[rb + g, t4 * 32 + m], O[rb + g, t4 * 32 + 16 + m] = d[lane][2], d[lane][3]
    return O.reshape(n_win, T, D)


@pytest.mark.parametrize("T", [1, 3, 7, 8])
def test_small_T_fragment_mapping_reproduces_attention(T):
    B = 3                                       # odd: the last warp handles a lone window
    g = torch.Generator().manual_seed(T)
    q = torch.randn(B, T, 128, generator=g, dtype=torch.float64)
    k = torch.randn(B, T, 128, generator=g, dtype=torch.float64)
    v = torch.randn(B, T, 128, generator=g, dtype=torch.float64)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    got = _attn_small_model(q.numpy(), k.numpy(), v.numpy(), lens.numpy(), T)
    s = (q @ k.transpose(1, 2)) / np.sqrt(128.0)
    for b, L in enumerate(lens.tolist()):
        s[b, :, L:] = -np.inf
    want = (torch.softmax(s, -1) @ v).numpy()
    assert np.abs(got - want).max() <= 1e-12


def test_logmel_fft_indexing_model():
    """csrc/k_logmel.cu: bit-reversed load + radix-2 DIT stages with the twiddle index j * (n/2 >> (s-1)),
    and the reflect-padded frame indexing, restated in NumPy."""
    def brev(i, bits):
        return int(format(i, f"0{bits}b")[::-1], 2)
    for n in (32, 512):
        log2n, half = n.bit_length() - 1, n // 2
        x = np.random.default_rng(n).standard_normal(n)
        tw = np.exp(-2j * np.pi * np.arange(half) / n)
        data = np.zeros(n, complex)
        for i in range(n):
            data[brev(i, log2n)] = x[i]
        for s in range(1, log2n + 1):
            hs, tstride = 1 << (s - 1), half >> (s - 1)
            for b in range(half):
                j = b & (hs - 1)
                i0 = ((b >> (s - 1)) << s) + j
                u, v = data[i0], data[i0 + hs] * tw[j * tstride]
                data[i0], data[i0 + hs] = u + v, u - v
        assert np.abs(data[:half + 1] - np.fft.rfft(x)).max() <= 1e-11
    n_samples, hop, n_fft = 1000, 160, 512
    a = np.arange(n_samples, dtype=np.float32)
    y = np.pad(a, n_fft // 2, mode="reflect")
    L = 1 + (len(y) - n_fft) // hop
    assert L == 1 + n_samples // hop
    for t in range(L):
        m = t * hop + np.arange(n_fft) - n_fft // 2
        m = np.where(m < 0, -m, m)
        m = np.where(m >= n_samples, 2 * (n_samples - 1) - m, m)
        np.testing.assert_array_equal(a[m], y[t * hop:t * hop + n_fft])
