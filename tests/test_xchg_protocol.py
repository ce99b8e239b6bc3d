"""CPU: the SyncBatchNorm peer-exchange protocol of csrc/xchg.cu restated with Python threads as ranks.

Each "rank" owns flags[slot][peer] and data[slot]; one exchange = publish into the own data slot, store the sequence
number into every peer's flag, spin on the own flags, add all peers' data in rank order, advance the own counter.  The
threads run many back-to-back exchanges with random stalls (a rank may race several exchanges ahead of a slow peer as
far as the protocol allows): every rank must read, for exchange k, exactly the values all ranks published for exchange
k — i.e. a slot is never overwritten while a peer may still read it (NSLOTS >= 2) — and all ranks must end with
identical sums (rank-ordered addition)."""
import random
import threading
import time

SLOTS = 4


def _run(world, rounds, seed):
    flags = [[[0] * world for _ in range(SLOTS)] for _ in range(world)]      # flags[owner][slot][peer]
    data = [[None] * SLOTS for _ in range(world)]                           # data[owner][slot]
    results = [[] for _ in range(world)]
    errors = []

    def rank_fn(r):
        rng = random.Random(seed * 100 + r)
        counter = 0
        for k in range(rounds):
            seq = counter + 1
            slot = seq % SLOTS
            if rng.random() < 0.2:
                time.sleep(rng.random() * 0.002)                            # rank skew
            data[r][slot] = (k, [float(r + 1) * (k + 1) + 0.125 * r])        # publish (tagged with the exchange id)
            for p in range(world):
                flags[p][slot][r] = seq                                      # st.release.sys into every peer's flags
            t0 = time.time()
            while any(flags[r][slot][p] != seq for p in range(world)):       # ld.acquire.sys spin on the own flags
                if time.time() - t0 > 5.0:
                    errors.append("rank %d stuck in exchange %d" % (r, k))
                    return
                time.sleep(0)
            acc = 0.0
            for p in range(world):                                           # rank-ordered sum over the peers' slots
                tag, vals = data[p][slot]
                if tag != k:
                    errors.append("rank %d read exchange %d data of rank %d during exchange %d" % (r, tag, p, k))
                    return
                acc += vals[0]
            results[r].append(acc)
            counter = seq

    threads = [threading.Thread(target=rank_fn, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return results, errors


def test_peer_exchange_protocol_is_race_free_and_rank_ordered():
    for world, seed in ((2, 1), (4, 2), (8, 3)):
        results, errors = _run(world, 300, seed)
        assert not errors, errors[:3]
        for r in range(1, world):
            assert results[r] == results[0]                                  # identical bits on every rank
        for k, v in enumerate(results[0]):
            expect = 0.0
            for p in range(world):
                expect += float(p + 1) * (k + 1) + 0.125 * p
            assert v == expect
