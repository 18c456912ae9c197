"""The synchronisation argument of the fused FedAvg kernel, checked on the index arithmetic alone (pure Python
mirror of csrc/fedavg.cu; no GPU): a per-CTA cross-GPU barrier is enough because tile t is packed and applied by
CTA (t div A) mod G on EVERY rank and reduced by the same CTA index on its owner rank (t mod A)."""
from hypothesis import given, settings, strategies as st


def pack_apply_tiles(block, n_tiles, A, G):
    """Tiles CTA `block` touches in phase 0 (pack) and phase 2 (apply): for q = block, block+G, ...; r < A."""
    out = []
    q = block
    while q * A < n_tiles:
        for r in range(A):
            t = q * A + r
            if t >= n_tiles:
                break
            out.append(t)
        q += G
    return out


def reduce_tiles(block, my_pos, n_tiles, A, G):
    """Tiles CTA `block` of the rank at position `my_pos` reduces + broadcasts in phase 1."""
    return list(range(my_pos + block * A, n_tiles, G * A))


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 16), st.integers(1, 40), st.integers(1, 5000))
def test_every_tile_has_one_owner_cta_and_the_same_cta_index_everywhere(A, G, n_tiles):
    packed = {}
    for b in range(G):
        for t in pack_apply_tiles(b, n_tiles, A, G):
            assert t not in packed, "tile packed twice"
            packed[t] = b
    assert sorted(packed) == list(range(n_tiles)), "pack / apply must cover the arena exactly once per rank"
    reduced = {}
    for pos in range(A):
        for b in range(G):
            for t in reduce_tiles(b, pos, n_tiles, A, G):
                assert t not in reduced, "tile reduced twice"
                reduced[t] = (pos, b)
    assert sorted(reduced) == list(range(n_tiles))
    for t in range(n_tiles):
        pos, b = reduced[t]
        assert pos == t % A                       # ownership
        assert b == packed[t] == (t // A) % G     # producer CTA == consumer CTA on every rank


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 16), st.sampled_from([8, 64, 148, 296]), st.integers(1, 3_000_000))
def test_adaptive_tile_size_gives_at_most_one_tile_per_rank_and_cta(A, G, n_blocks):
    n = n_blocks * 2048                            # ParamArena pads to 2048 elements
    per = -(-n // (A * G))
    tile = max(1024, (per + 31) // 32 * 32)        # FedAvgSession.aggregate
    assert tile % 32 == 0                          # fp8 wire: 32-element scale blocks never straddle tiles
    n_tiles = -(-n // tile)
    assert n_tiles <= A * G or tile == 1024        # one tile per (live rank, CTA) unless the arena is tiny
    for b in range(G):
        assert len(pack_apply_tiles(b, n_tiles, A, G)) <= max(A, -(-n_tiles // G))
