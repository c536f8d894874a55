"""CPU twin of the work partition of the stream-form decode attention (csrc/llm_attention.cu: llm_attn_decode_stream_kernel).
The kernel flattens (sequence, kv head, 64-key block) in that order, cuts the list into one contiguous range per CTA and lets
the parts of a split (sequence, kv head) find each other through closed-form arithmetic (lds_owner, the two workspace slots
per CTA, the arrival count).  These tests restate that arithmetic in Python and check its invariants over many shapes --
the GPU tests then only have to show that the kernel computes the right numbers on a few of them."""
import random

import pytest


def partition(ctx_lens, kvh, grid, max_ctx=1 << 30):
    """what every CTA computes for itself: prefix, T, n_cta, [f0, f1)"""
    nb = [(min(c, max_ctx - 1) + 64) // 64 for c in ctx_lens]          # blocks of a sequence that holds positions 0 .. pos
    prefix = [0]
    for n in nb:
        prefix.append(prefix[-1] + n)
    T = prefix[-1] * kvh
    n_cta = min(grid, T)
    ranges = [((T * c) // n_cta, (T * (c + 1)) // n_cta) for c in range(n_cta)]
    return nb, prefix, T, n_cta, ranges


def owner(f, T, n):
    return ((f + 1) * n - 1) // T                                       # lds_owner


def seek(prefix, kvh, n_seq, f):
    lo, hi = 0, n_seq - 1                                               # lds_seek: binary search over the block prefix
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if prefix[mid] * kvh <= f:
            lo = mid
        else:
            hi = mid - 1
    nbb = prefix[lo + 1] - prefix[lo]
    rem = f - prefix[lo] * kvh
    return lo, rem // nbb, rem % nbb, nbb


def segments_of(cta_range, prefix, kvh, n_seq):
    """the (b, h, j0, j1, nb, seg0) runs a CTA walks, as the consumer loop forms them"""
    f, f1 = cta_range
    out = []
    while f < f1:
        b, h, j0, nbb = seek(prefix, kvh, n_seq, f)
        seg0 = f - j0
        j1 = min(nbb, f1 - seg0)
        out.append((b, h, j0, j1, nbb, seg0))
        f += j1 - j0
    return out


CASES = [([512] * 32, 8, 296), ([512] * 32, 4, 296), ([5], 2, 296), ([700, 0, 62, 63, 64, 126, 127, 128, 299, 510, 511, 639], 2, 296),
         ([3000] * 4, 8, 296), ([6000], 8, 148), ([1] * 32, 8, 296), ([1023] * 32, 8, 148)]
for seed in range(40):
    r = random.Random(seed)
    CASES.append(([r.randrange(0, 2000) for _ in range(r.randrange(1, 33))], r.choice([1, 2, 4, 8]), r.choice([148, 296, 7, 64])))


@pytest.mark.parametrize("ctx_lens,kvh,grid", CASES)
def test_every_block_is_owned_once_and_parts_find_each_other(ctx_lens, kvh, grid):
    n_seq = len(ctx_lens)
    nb, prefix, T, n_cta, ranges = partition(ctx_lens, kvh, grid)
    assert all(f1 > f0 for f0, f1 in ranges), "every CTA of the partition owns at least one block"
    assert ranges[0][0] == 0 and ranges[-1][1] == T and all(ranges[i][1] == ranges[i + 1][0] for i in range(n_cta - 1))
    # the closed form names the CTA whose range holds a block
    for f in set([0, T - 1] + [random.Random(1).randrange(T) for _ in range(50)] + [r[0] for r in ranges]):
        c = owner(f, T, n_cta)
        assert ranges[c][0] <= f < ranges[c][1]
    covered = {}
    arrivals = {}
    slots_used = set()
    for c, rg in enumerate(ranges):
        pend = 0
        for (b, h, j0, j1, nbb, seg0) in segments_of(rg, prefix, kvh, n_seq):
            assert nbb == nb[b] and 0 <= j0 < j1 <= nbb
            for j in range(j0, j1):
                assert (b, h, j) not in covered
                covered[(b, h, j)] = c
            whole = j0 == 0 and j1 == nbb
            if whole:
                continue
            c_first = owner(seg0, T, n_cta)
            parts = owner(seg0 + nbb - 1, T, n_cta) - c_first + 1
            part = c - c_first
            assert parts >= 2 and 0 <= part < parts
            # the merging CTA looks for part p in CTA c_first + p, slot (p == 0 ? 1 : 0): the parts sit in consecutive CTAs
            slot = 1 if part == 0 else 0
            assert (c, slot) not in slots_used, "a CTA's head and tail partials must not share a workspace slot"
            slots_used.add((c, slot))
            arrivals.setdefault((b, h), []).append((part, parts, c))
            pend += 1
        assert pend <= 2, "a range has at most two split segments (its head and its tail)"
    assert len(covered) == T and all((b, h, j) in covered for b in range(n_seq) for h in range(kvh) for j in range(nb[b]))
    for key, lst in arrivals.items():
        parts = lst[0][1]
        assert sorted(p for p, _, _ in lst) == list(range(parts)), "the arrival count equals the number of parts, each part once"
        cs = [c for _, _, c in sorted(lst)]
        assert cs == list(range(cs[0], cs[0] + parts))


def test_new_row_lives_in_the_last_block():
    for pos in (0, 1, 62, 63, 64, 65, 127, 128, 1000):
        nb = (pos + 64) // 64
        assert pos >> 6 == nb - 1
