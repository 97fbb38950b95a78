"""not gpu: the N>1 plumbing (ebwt2snp_b200/sharding.py) over a real process group -- gloo, world sizes 2 and 3,
one process per rank on CPU.  Every rank emulates what its GPU shard would report (numpy restatement of the
stencil, tests/helpers.py) and the ranks must assemble exactly the oracle's .clusters / statistics / event ids."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ebwt2snp_b200 import api, sharding
        from oracle import oracle as O
        from tests import helpers as H

        rng = np.random.default_rng(seed)  # same stream on every rank: same global arrays
        dev = torch.device("cpu")
        for it in range(30):
            n = int(rng.integers(400, 6000))
            k = int(rng.choice([2, 5, 16]))
            m = int(rng.choice([1, 2, 3]))
            lcp = H.random_lcp(rng, n, k, it % 5)
            bwt = rng.choice(H.BWT_ALPHABET, size=n)
            if it % 3 == 0:
                cuts = sharding.shard_cuts(n, world)
            else:  # uneven cuts, possibly inside clusters
                inner = sorted(int(c) for c in rng.choice(np.arange(2, n - 2), size=world - 1, replace=False))
                cuts = [0] + inner + [n]
                if any(b - a < 2 for a, b in zip(cuts[:-1], cuts[1:])):
                    cuts = sharding.shard_cuts(n, world)
            lo, hi = cuts[rank], cuts[rank + 1]
            summary, rs, rl = H.emulate_shard(lcp, bwt, lo, hi, k, m)

            # ---- step 2: summaries -> merge ----
            mg, sums = sharding.merge_clusters(summary, dev)
            assert len(sums) == world and sums[rank].global_off == lo
            S, L = [], []
            if mg.n_prepend and mg.prepend_written:
                S.append(mg.prepend_start)
                L.append(mg.prepend_len)
            S += rs.tolist()
            L += rl.tolist()
            for i in range(mg.n_append):
                S.append(mg.append_start[i])
                L.append(mg.append_len[i])
            assert mg.record_offset + len(S) <= mg.total_written

            # ---- step 3: statistics ----
            st = api.Stats()
            for v in L:
                if v <= api.MAX_C_LEN:
                    st.hist[v] += 1
            st.n_clust, st.n_bases = len(L), int(sum(L))
            st.last_len = L[-1] if L else 0
            es, el, enc, _ = O.cluster_lm(lcp, bwt, k, m)
            if len(es):
                tot = sharding.merge_statistics(st, 2, 0.9, dev)
                ost = O.statistics(es, el, 2, 0.9)
                assert list(tot.hist) == list(ost.hist)
                assert (tot.n_clust, tot.n_bases, tot.max_clust_length) == (ost.n_clust, ost.n_bases, ost.max_clust_length)

            # ---- steps 2 + 3 in ONE collective (what hot_path_step does): own-record histogram in, merged view + global stats out ----
            own = api.Stats()
            for v in rl.tolist():
                if v <= api.MAX_C_LEN:
                    own.hist[v] += 1
            own.n_clust, own.n_bases = len(rl), int(rl.astype(np.int64).sum())
            own.last_len = int(rl[-1]) if len(rl) else 0
            if len(es):
                mg2, tot2 = sharding.exchange_and_merge(summary, own, 2, 0.9, dev)
                assert bytes(mg2) == bytes(mg)
                ost = O.statistics(es, el, 2, 0.9)
                assert list(tot2.hist) == list(ost.hist), (it, cuts)
                assert (tot2.n_clust, tot2.n_bases, tot2.max_clust_length) == (ost.n_clust, ost.n_bases, ost.max_clust_length), (it, cuts)

            # ---- step 4: event ids ----
            fake_events = int(rng.integers(0, 50)) + rank  # differs per rank
            first, total = sharding.event_id_offset(fake_events, dev)
            counts = [g[0] for g in sharding.all_gather_words([fake_events], dev)]
            assert first == 1 + sum(counts[:rank]) and total == sum(counts)

            # ---- global .clusters = concatenation in rank order ----
            gathered = [None] * world
            dist.all_gather_object(gathered, (mg.record_offset, S, L, mg.n_clust_out & 0xFFFFFFFF, mg.total_written))
            if rank == 0:
                allS, allL = [], []
                for off, s_, l_, ncl, totw in gathered:
                    assert off == len(allS) and ncl == enc and totw == len(es)
                    allS += s_
                    allL += l_
                assert allS == es.tolist() and allL == el.tolist(), (it, cuts)

        # ---- step 1: halo exchange ----
        n = 1000
        base = rank * n
        lcp_t = torch.arange(base, base + n, dtype=torch.int32)
        left, right = sharding.exchange_halo(lcp_t, lcp_t + 1, lcp_t + 2, (lcp_t % 200).to(torch.uint8), dev)
        if rank > 0:
            assert left["lcp"].tolist() == [base - 2, base - 1] and left["suff"].tolist() == [base, base + 1]
        else:
            assert left is None
        if rank < world - 1:
            assert right["lcp"].tolist() == list(range(base + n, base + n + sharding.HALO_R))
            assert right["bwt"].tolist() == [(base + n + i) % 200 for i in range(sharding.HALO_R)]
        else:
            assert right is None
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, "FAIL: " + "".join(traceback.format_exception(e))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_over_gloo(built, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 77 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in sorted(res):
        assert msg == "ok", f"rank {rank}: {msg}"


def test_shard_cuts():
    from ebwt2snp_b200 import sharding
    for n in (2, 3, 5, 17, 1000, 12345):
        for parts in (1, 2, 3, 8):
            c = sharding.shard_cuts(n, parts)
            assert c[0] == 0 and c[-1] == n and all(b - a >= 2 for a, b in zip(c[:-1], c[1:]))
            assert max(b - a for a, b in zip(c[:-1], c[1:])) - min(b - a for a, b in zip(c[:-1], c[1:])) <= 1
