"""CPU: restatement of the residual-tile prefetch protocol of csrc/gemm_fused.cu (one epilogue warp).

Per valid 32-column chunk q the warp (1) issues the TMA load of the NEXT valid chunk into buffer (q + 1) & 1 after
waiting for the output store of chunk q - 1 to have read that buffer, (2) waits for chunk q's own tile on buffer q & 1
with a phase bit that flips per use, (3) reuses buffer q & 1 as the output staging and TMA-stores it.  The simulation
checks, over ragged shapes, that every wait has exactly one matching load (same tile coordinates, same buffer, same
phase), that no load ever lands in a buffer whose previous contents are still needed, and that nothing is left in
flight at the end."""
import itertools

BM, BN, CPW = 128, 128, 2


def simulate(M, N, grid, block, warp):
    tiles_n = (N + BN - 1) // BN
    num_tiles = ((M + BM - 1) // BM) * tiles_n
    quarter, col_w0 = warp & 3, (warp >> 2) * (CPW * 32)

    def valid(tile, cl):
        return tile < num_tiles and (tile % tiles_n) * BN + col_w0 + cl * 32 < N

    def next_chunk(tile, cl):
        while True:
            cl += 1
            if cl == CPW:
                cl, tile = 0, tile + grid
            if tile >= num_tiles or valid(tile, cl):
                return tile, cl

    buf_state = [None, None]      # None = free; ("load", coords, phase) | ("store", coords)
    completed = [0, 0]            # loads completed per buffer (mbarrier phase = completed & 1 once waited)
    rphase = [0, 0]
    q = 0
    processed = []

    def coords(tile, cl):
        return ((tile // tiles_n) * BM + quarter * 32, (tile % tiles_n) * BN + col_w0 + cl * 32)

    def issue(tile, cl, buf):
        # the previous user of `buf` must be done: either never used, or its output store has been waited for
        assert buf_state[buf] is None, ("load into a busy buffer", buf, buf_state[buf])
        buf_state[buf] = ("load", coords(tile, cl), completed[buf] & 1)
        completed[buf] += 1

    t0, c0 = next_chunk(block, -1)
    if t0 < num_tiles:
        issue(t0, c0, 0)
    tile = block
    while tile < num_tiles:
        for cl in range(CPW):
            if not valid(tile, cl):
                continue
            buf = q & 1
            nt, nc = next_chunk(tile, cl)
            if nt < num_tiles:
                # tma_store_wait_read(): every earlier output store has finished reading shared memory
                for b in (0, 1):
                    if buf_state[b] is not None and buf_state[b][0] == "store":
                        buf_state[b] = None
                issue(nt, nc, buf ^ 1)
            st = buf_state[buf]
            assert st is not None and st[0] == "load" and st[1] == coords(tile, cl), (tile, cl, buf, st)
            assert st[2] == rphase[buf], "mbarrier phase mismatch"
            rphase[buf] ^= 1
            buf_state[buf] = ("store", coords(tile, cl))      # same buffer becomes the output staging
            processed.append((tile, cl))
            q += 1
        tile += grid
    assert all(s is None or s[0] == "store" for s in buf_state), "a residual load was never consumed"
    return processed


def test_prefetch_protocol_over_ragged_shapes():
    total = 0
    for M, N, grid in itertools.product((100, 128, 1000, 5000), (72, 128, 136, 256, 520), (1, 3, 7, 296)):
        tiles = ((M + BM - 1) // BM) * ((N + BN - 1) // BN)
        g = min(grid, tiles)
        seen = set()
        for block in range(g):
            for warp in range(8):
                for t, c in simulate(M, N, g, block, warp):
                    key = (t, warp, c)
                    assert key not in seen
                    seen.add(key)
                    total += 1
        # every valid (tile, column chunk) is handled exactly once by each of the 4 row-quarter warps of its column group
        expect = sum(4 for t in range(tiles) for grp in range(2) for c in range(CPW)
                     if (t % ((N + BN - 1) // BN)) * BN + grp * 64 + c * 32 < N)
        assert len(seen) == expect, (M, N, grid, len(seen), expect)
    assert total > 1000
