"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun, NCCL).  Shards one seeded read set by contiguous
eBWT ranges (deliberately uneven cuts), runs the hot path with ebwt2snp_b200.sharding and checks on rank 0 that the
concatenated .clusters records and the .snp text equal the oracle's single-pass result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ebwt2snp_b200 import api, sharding, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    k, m = 16, 2
    for seed, name in ((3, "tiny"), (4, "small")):
        rs = synth.make_config(name, seed=seed)          # same data on every rank (seeded)
        e = synth.build_egsa(rs.reads)
        eg = {kk: (v.numpy() if hasattr(v, "numpy") else v) for kk, v in e.items()}
        for f in ("lcp", "text", "suff"):
            eg[f] = eg[f].view(np.uint32)
        n = int(eg["n"])
        rng = np.random.default_rng(100 + seed)
        cuts = [0] + sorted(int(c) for c in rng.choice(np.arange(1000, n - 1000), size=world - 1, replace=False)) + [n]
        lo, hi = cuts[rank], cuts[rank + 1]
        ctx = api.Context(local)
        sh = ctx.shard(hi - lo, lo, n)
        a, b = max(0, lo - sharding.HALO_L), min(n, hi + sharding.HALO_R)
        sh.load_soa(eg["lcp"][a:b], eg["text"][a:b], eg["suff"][a:b], eg["bwt"][a:b], first=a)
        sh.seal()
        off = O.uniform_read_offsets(*rs.reads.shape)
        ctx.stage_reads(rs.reads, off)
        p = api.default_params(rs.nreads1)
        # once through the torch.distributed exchange, once through the library's own NCCL communicator: same result
        mg, st, cnt, ids = sharding.hot_path_step(sh, p, k, m, dev)
        comm = sharding.make_comm(ctx, dev)
        mg2, st2, cnt2, _ = sharding.hot_path_step(sh, p, k, m, dev, comm=comm)
        assert bytes(mg2) == bytes(mg) and bytes(st2) == bytes(st) and bytes(cnt2) == bytes(cnt), "native exchange differs"
        # ... and end to end from host records: every rank streams its range through a chunked shard (e2s_pipeline_host_sharded)
        a2, b2 = max(0, lo - 176), min(n, hi + 152)
        sub = {kk: eg[kk][a2:b2] for kk in ("lcp", "text", "suff", "bwt")}
        sub["n"] = b2 - a2
        rec = synth.gesa_records(sub).view(np.uint8).reshape(-1)
        os.environ["E2S_CHUNK_POSITIONS"] = str(32768 if name == "tiny" else 1 << 20)
        ctx2 = api.Context(local)
        comm2 = sharding.make_comm(ctx2, dev)
        rec10 = np.empty((len(sh.cluster_fetch_packed()) // 10 + 16) * 10, dtype=np.uint8)
        evbuf = (api.Event * (int(cnt.n_variants) + 16))()
        res3, mg3, st3, m3 = api.pipeline_host_sharded(ctx2, comm2, rec, a2, lo, hi - lo, n, rs.reads.reshape(-1), off, p, k, m,
                                                       rec10=rec10, events=evbuf)
        assert bytes(mg3) == bytes(mg) and bytes(st3) == bytes(st), "chunked sharded pipeline: merge / statistics differ"
        assert rec10[: m3 * 10].tobytes() == sh.cluster_fetch_packed(), "chunked sharded pipeline: records differ"
        assert (res3.snp.n_candidates, res3.snp.n_events) == (cnt.n_candidates, cnt.n_events)
        assert api.events_format(list(evbuf)[: res3.snp.n_variants], p) == api.events_format(sh.events(), p)
        comm2.close()
        ctx2.close()
        comm.close()
        first_id, total_events = ids.resolve()
        recs = sh.cluster_fetch_packed()
        text = api.events_format(sh.events(), p, first_id=first_id)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mg.record_offset, recs, text, int(cnt.n_candidates), int(mg.n_clust_out), int(st.max_clust_length)))
        if rank == 0:
            es, el, enc, _ = O.cluster_lm(eg["lcp"], eg["bwt"], k, m)
            op = O.default_params(rs.nreads1)
            ost = O.statistics(es, el, op.mcov_out, op.pval)
            otext, ores = O.find_events(eg["lcp"], eg["text"], eg["suff"], eg["bwt"], es, el, op, ost.max_clust_length, rs.reads, off)
            all_recs = b"".join(g[1] for g in gathered)
            assert [g[0] for g in gathered] == list(np.cumsum([0] + [len(g[1]) // 10 for g in gathered[:-1]])), "record offsets"
            assert all_recs == O.clusters_to_bytes(es, el), f"{name}: .clusters differ"
            assert gathered[-1][4] & 0xFFFFFFFF == enc
            assert all(g[5] == ost.max_clust_length for g in gathered)
            assert sum(g[3] for g in gathered) == ores.n_candidates
            assert b"".join(g[2] for g in gathered) == otext, f"{name}: .snp differ"
            print(f"OK {name} world={world} cuts={cuts} records={len(all_recs) // 10} events={otext.count(b'>') // 2}", flush=True)
        sh.close()
        ctx.close()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
