"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference binaries
(oracle/_ref, built from /root/reference by oracle/Makefile) on seeded inputs.

    python tests/golden/make_golden.py

Only runs where oracle/_ref exists (the build container).  The fixtures hold the inputs in compact form
(2-bit packed reads + generator parameters; the EGSA is rebuilt deterministically by
ebwt2snp_b200.synth and checked against the recorded sha256 of the .gesa the reference consumed) and the
reference's own outputs (.clusters records, .snp text, printed counts)."""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ebwt2snp_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GESA_DT = np.dtype([("text", "<u4"), ("suff", "<u4"), ("lcp", "<u4"), ("bwt", "u1")])

MICRO = {
    "micro_a": dict(gen=dict(G=6000, reads_per_sample=1200, L=100, n_snps=12, n_indels=3, rc=True, seed=11), k=16, m=2),
    "micro_b": dict(gen=dict(G=4000, reads_per_sample=1500, L=64, n_snps=10, n_indels=2, rc=False, seed=12), k=12, m=3),
}
# index layouts beyond the 4/4/4 .gesa of the main fixtures: (x, y, z, bcr)
LAYOUTS = [(1, 4, 1, False), (2, 4, 2, False), (4, 4, 4, True), (1, 4, 1, True), (8, 8, 8, False)]
SNP_VARIANTS = [
    [], ["-m", "3"], ["-c", "3", "-g", "4"], ["-L", "25", "-R", "20", "-e", "1"], ["-p", "0.5"], ["-v", "0"],
]


def pack2(reads):
    code = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), reads).astype(np.uint8)
    flat = code.reshape(-1)
    pad = (-len(flat)) % 4
    flat = np.concatenate([flat, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (flat[:, 0] | (flat[:, 1] << 2) | (flat[:, 2] << 4) | (flat[:, 3] << 6)).astype(np.uint8)


def phase1_fuzz(d):
    rng = np.random.default_rng(2024)
    fa = os.path.join(d, "F.fasta")
    open(fa, "w").write(">a\nA\n")
    alphabet = np.frombuffer(b"ACGT$acgt\x00\xff", dtype=np.uint8)
    lcps, bwts, meta, outs = [], [], [], []
    for it in range(72):
        n = int(rng.integers(2, 40)) if it % 3 == 0 else int(rng.integers(40, 2500))
        k = int(rng.choice([1, 2, 3, 5, 16, 70]))
        m = int(rng.choice([1, 2, 3, 8]))
        mode = it % 4
        if mode == 0:
            lcp = rng.integers(0, 2 * k + 2, size=n)
        elif mode == 1:
            lcp = rng.integers(0, 70000, size=n)
        elif mode == 2:
            lcp = np.maximum(0, np.cumsum(rng.integers(-3, 4, size=n)) + k)
        else:
            lcp = rng.integers(k, k + 3, size=n)
        lcp = lcp.astype(np.uint32)
        bwt = rng.choice(alphabet, size=n)
        rec = np.zeros(n, dtype=GESA_DT)
        rec["lcp"], rec["bwt"] = lcp, bwt
        rec["text"] = rng.integers(0, 1000, size=n)
        rec["suff"] = rng.integers(0, 100, size=n)
        rec.tofile(fa + ".gesa")
        r, ncl = O.ref_ebwt2clust(fa, k=k, m=m)
        assert r.returncode == 0
        out = np.frombuffer(open(fa + ".clusters", "rb").read(), dtype=np.uint8)
        lcps.append(lcp)
        bwts.append(bwt)
        outs.append(out)
        meta.append((n, k, m, ncl, len(out)))
    np.savez_compressed(os.path.join(HERE, "phase1_fuzz.npz"), lcp=np.concatenate(lcps), bwt=np.concatenate(bwts),
                        out=np.concatenate(outs), meta=np.array(meta, dtype=np.int64))


def phase1_fuzz_widths(d):
    """ebwt2clust with -x 1 / 2 / 8 (and narrow y, z): the post-EOF phantom LCP is truncated to the field width."""
    rng = np.random.default_rng(4048)
    fa = os.path.join(d, "W.fasta")
    open(fa, "w").write(">a\nA\n")
    alphabet = np.frombuffer(b"ACGT$acgt\x00\xff", dtype=np.uint8)
    lcps, bwts, meta, outs = [], [], [], []
    combos = [(1, 4, 1), (1, 4, 4), (2, 4, 4), (8, 4, 4), (1, 1, 1), (2, 2, 2), (8, 8, 8), (4, 4, 1)]
    for it in range(96):
        x, y, z = combos[it % len(combos)]
        n = int(rng.integers(3, 60)) if it % 3 == 0 else int(rng.integers(60, 1200))
        k = int(rng.choice([1, 2, 3, 5, 16, 70]))
        m = int(rng.choice([1, 2, 3, 8]))
        mode = (it // len(combos)) % 4
        if mode == 0:
            lcp = rng.integers(0, 2 * k + 2, size=n)
        elif mode == 1:
            lcp = rng.integers(0, 70000, size=n)
        elif mode == 2:
            lcp = np.maximum(0, np.cumsum(rng.integers(-3, 4, size=n)) + k)
        else:
            lcp = rng.integers(k, k + 3, size=n)
        lcp = np.minimum(lcp, (1 << (8 * min(x, 4))) - 1).astype(np.uint32)
        bwt = rng.choice(alphabet, size=n)
        text = rng.integers(0, 200, size=n)
        suff = rng.integers(0, 100, size=n)
        open(fa + ".gesa", "wb").write(_gesa_bytes(lcp, text, suff, bwt, x, y, z))
        r, ncl = O.ref_ebwt2clust(fa, k=k, m=m, x=x, y=y, z=z)
        assert r.returncode == 0
        out = np.frombuffer(open(fa + ".clusters", "rb").read(), dtype=np.uint8)
        lcps.append(lcp)
        bwts.append(bwt)
        outs.append(out)
        meta.append((n, k, m, ncl, len(out), x, y, z))
    np.savez_compressed(os.path.join(HERE, "phase1_fuzz_widths.npz"), lcp=np.concatenate(lcps), bwt=np.concatenate(bwts),
                        out=np.concatenate(outs), meta=np.array(meta, dtype=np.int64))


def micro(name, cfg, d):
    rs = synth.make_read_set(**cfg["gen"])
    e = synth.build_egsa(rs.reads)
    fa = synth.write_dataset(d, rs, e, name=name + ".fasta")
    sha = hashlib.sha256(open(fa + ".gesa", "rb").read()).hexdigest()
    r, ncl = O.ref_ebwt2clust(fa, k=cfg["k"], m=cfg["m"])
    assert r.returncode == 0
    clusters = np.frombuffer(open(fa + ".clusters", "rb").read(), dtype=np.uint8)
    snp_path = os.path.join(d, name + ".snp")
    res = {}
    for vi, extra in enumerate(SNP_VARIANTS):
        if os.path.exists(snp_path):
            os.remove(snp_path)
        r, info = O.ref_clust2snp(fa, rs.nreads1, extra=extra)
        res[f"v{vi}_args"] = np.array(" ".join(extra))
        res[f"v{vi}_rc"] = np.int64(info["returncode"])
        res[f"v{vi}_allowed"] = np.array(info.get("allowed", (-1, -1)), dtype=np.int64)
        res[f"v{vi}_ncand"] = np.int64(info.get("n_candidates", -1))
        res[f"v{vi}_snp"] = np.frombuffer(open(snp_path, "rb").read() if os.path.exists(snp_path) else b"", dtype=np.uint8)
    # the same read set through other index layouts (narrow fields, BCR triple): the phantom record differs
    for j, (x, y, z, bcr) in enumerate(LAYOUTS):
        d2 = os.path.join(d, f"lay{j}")
        fa2 = synth.write_dataset(d2, rs, e, name=name + ".fasta", x=x, y=y, z=z, bcr=bcr)
        r, ncl2 = O.ref_ebwt2clust(fa2, k=cfg["k"], m=cfg["m"], x=x, y=y, z=z)
        assert r.returncode == 0
        r, info = O.ref_clust2snp(fa2, rs.nreads1, x=x, y=y, z=z, timeout=600)
        sp = os.path.join(d2, name + ".snp")
        res[f"lay{j}_spec"] = np.array([x, y, z, int(bcr)], dtype=np.int64)
        res[f"lay{j}_nclust"] = np.int64(ncl2)
        res[f"lay{j}_clusters"] = np.frombuffer(open(fa2 + ".clusters", "rb").read(), dtype=np.uint8)
        res[f"lay{j}_rc"] = np.int64(info["returncode"])
        res[f"lay{j}_allowed"] = np.array(info.get("allowed", (-1, -1)), dtype=np.int64)
        res[f"lay{j}_ncand"] = np.int64(info.get("n_candidates", -1))
        res[f"lay{j}_snp"] = np.frombuffer(open(sp, "rb").read() if os.path.exists(sp) else b"", dtype=np.uint8)
    res["n_layouts"] = np.int64(len(LAYOUTS))
    gen = cfg["gen"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), reads2bit=pack2(rs.reads), shape=np.array(rs.reads.shape),
                        nreads1=np.int64(rs.nreads1), gen=np.array(repr(gen)), k=np.int64(cfg["k"]), m=np.int64(cfg["m"]),
                        gesa_sha256=np.array(sha), n_clust_out=np.int64(ncl), clusters=clusters, n_variants=np.int64(len(SNP_VARIANTS)),
                        **res)
    print(name, "n =", e["n"], "clusters =", len(clusters) // 10, "closed =", ncl,
          "events(v0) =", bytes(res["v0_snp"]).count(b">") // 2)


def _gesa_bytes(lcp, text, suff, bwt, x, y, z):
    n = len(lcp)
    out = np.zeros((n, x + y + z + 1), dtype=np.uint8)

    def put(col, arr, w):
        a = np.asarray(arr, dtype=np.uint64)
        for b in range(w):
            out[:, col + b] = (a >> np.uint64(8 * b)) & np.uint64(0xFF)
    put(0, text, y)
    put(y, suff, z)
    put(y + z, lcp, x)
    out[:, y + z + x] = bwt
    return out.tobytes()


def phantom_tail(d):
    """Hand-made index whose LAST cluster reaches position n (the post-EOF phantom record) and passes the
    find_variants filters only if the phantom record lands in sample 1: counts[0]['A'] = 5, counts[1]['C'] = 4 (+1).
    The number of candidates the reference prints for a few -n values reads out the phantom's `text` field."""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref")
    fa = os.path.join(d, "T.fasta")
    open(fa, "w").write(">a\nACGT\n")
    combos = [(4, 4, 4), (1, 4, 1), (1, 4, 4), (2, 4, 2), (4, 4, 1), (8, 8, 8), (1, 1, 1), (2, 2, 2), (1, 2, 4), (4, 2, 1), (2, 4, 4)]
    rows, arrays = [], {}
    ci = 0
    for bcr in (False, True):
        for (x, y, z) in combos:
            mk = lambda w: (1 << (8 * min(w, 4))) - 1
            lcpv, sufv, txtv = 0x11223344 & mk(x), 0x55667788 & mk(z), 0x0003BBCC & mk(y)
            n = 40
            V = lcpv if lcpv >= 16 else 40
            lcp = np.zeros(n, dtype=np.uint64)
            text = np.zeros(n, dtype=np.uint64)
            suff = np.full(n, 50 & mk(z), dtype=np.uint64)
            bwt = np.full(n, ord("G"), dtype=np.uint8)
            lcp[n - 9:] = V
            bwt[n - 9:n - 4] = ord("A")
            bwt[n - 4:] = ord("C")
            text[n - 4:] = txtv
            suff[n - 4:] = 3
            suff[n - 1] = sufv
            for suffix in (".gesa", ".out", ".out.lcp", ".out.pairSA", ".clusters"):
                if os.path.exists(fa + suffix):
                    os.remove(fa + suffix)
            if bcr:
                bwt.tofile(fa + ".out")
                lcp.astype(f"<u{x}").tofile(fa + ".out.lcp")
                rec = np.empty(n, dtype=np.dtype([("s", f"<u{z}"), ("t", f"<u{y}")]))
                rec["s"], rec["t"] = suff, text
                rec.tofile(fa + ".out.pairSA")
            else:
                open(fa + ".gesa", "wb").write(_gesa_bytes(lcp, text, suff, bwt, x, y, z))
            r = subprocess.run([os.path.join(ref, "ebwt2clust"), "-i", fa, "-x", str(x), "-y", str(y), "-z", str(z)],
                               capture_output=True, text=True, timeout=60)
            assert r.returncode == 0
            clusters = np.frombuffer(open(fa + ".clusters", "rb").read(), dtype=np.uint8)
            slot = O.lib().oracle_phantom_slot  # the model under test: only used to choose informative -n values
            slot.restype = __import__("ctypes").c_uint64
            sv = slot(int(lcp[-1]) & 0xFFFFFFFF, int(text[-1]) & 0xFFFFFFFF, int(suff[-1]) & 0xFFFFFFFF, int(bwt[-1]), x, y, z, int(bcr))
            ptext = sv & mk(y)
            probes = sorted({1, max(1, ptext - 1), max(1, ptext), ptext + 1, max(1, txtv), txtv + 1})
            for n1 in probes:
                cmd = [os.path.join(ref, "clust2snp"), "-i", fa, "-n", str(n1), "-x", str(x), "-y", str(y), "-z", str(z), "-L", "1", "-R", "1"]
                try:
                    out = subprocess.run(cmd, capture_output=True, text=True, timeout=5).stdout
                except subprocess.TimeoutExpired as ex:  # the reference goes on to build a huge read table: the count is already printed
                    out = ex.stdout.decode() if isinstance(ex.stdout, bytes) else (ex.stdout or "")
                nc = -1
                for line in out.splitlines():
                    if line.startswith("Done. ") and "potential" in line:
                        nc = int(line.split()[1])
                assert nc >= 0, (x, y, z, bcr, n1)
                rows.append((ci, x, y, z, int(bcr), n1, nc))
            arrays[f"c{ci}_lcp"], arrays[f"c{ci}_text"], arrays[f"c{ci}_suff"], arrays[f"c{ci}_bwt"] = lcp, text, suff, bwt
            arrays[f"c{ci}_clusters"] = clusters
            ci += 1
    np.savez_compressed(os.path.join(HERE, "phantom_tail.npz"), rows=np.array(rows, dtype=np.int64), n_cases=np.int64(ci), **arrays)
    print("phantom_tail:", ci, "cases,", len(rows), "readouts; candidates seen:", sorted({r[-1] for r in rows}))


def snp_text_tools(d):
    """filter_snp / snp2fastq of the reference on the fixtures' .snp files and on damaged copies of them
    (missing fields keep the previous group's value in the reference: its variables persist)."""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref")
    rng = np.random.default_rng(77)
    inputs = []
    for name in MICRO:
        z = np.load(os.path.join(HERE, name + ".npz"))
        inputs.append(bytes(z["v0_snp"]))
        inputs.append(bytes(z["v2_snp"]))
    base = inputs[0].decode().split("\n")
    for _ in range(6):  # damaged variants: cut bars / underscores / colons out of some header lines, drop trailing newline
        lines = list(base[:4 * int(rng.integers(3, 12))])
        for i in range(0, len(lines), 2):
            if rng.random() < 0.4:
                ch = str(rng.choice(["|", "_", ":"]))
                parts = lines[i].split(ch)
                if len(parts) > 2:
                    cut = int(rng.integers(1, len(parts)))
                    lines[i] = ch.join(parts[:cut])
            if rng.random() < 0.15:
                lines[i] += "|"
        text = "\n".join(lines) + ("\n" if rng.random() < 0.5 else "")
        inputs.append(text.encode())
    out = {"n": np.int64(len(inputs))}
    for j, data in enumerate(inputs):
        path = os.path.join(d, f"t{j}.snp")
        open(path, "wb").write(data)
        out[f"in{j}"] = np.frombuffer(data, dtype=np.uint8)
        for M in (0, 5, 9):
            r = subprocess.run([os.path.join(ref, "filter_snp"), path, str(M)], capture_output=True, timeout=60)
            assert r.returncode == 0
            out[f"filter{j}_{M}"] = np.frombuffer(r.stdout, dtype=np.uint8)
        for flag in ([], ["-i"]):
            if os.path.exists(path + ".fastq"):
                os.remove(path + ".fastq")
            r = subprocess.run([os.path.join(ref, "snp2fastq"), path, *flag], capture_output=True, timeout=60)
            assert r.returncode == 0
            out[f"fastq{j}_{len(flag)}"] = np.frombuffer(open(path + ".fastq", "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "snp_text.npz"), **out)
    print("snp_text:", len(inputs), "inputs")


def snp_vs_vcf_inputs():
    """Planted-truth inputs for snp_vs_vcf: the generator knows the SNPs it planted, so it can
    write the VCF and the sample-1 reference; the calls are the .snp text of the oracle's ebwt2clust + clust2snp on the
    same read set.  -> list of (name, {fasta, vcf, calls: bytes}, extra argv)"""
    from ebwt2snp_b200 import synth
    rs = synth.make_config("tiny", seed=21)
    e = synth.build_egsa(rs.reads)
    lcp, text, suff = (e[k].numpy().view(np.uint32) for k in ("lcp", "text", "suff"))
    bwt = e["bwt"].numpy()
    es, el, _, _ = O.cluster_lm(lcp, bwt, 16, 2)
    op = O.default_params(rs.nreads1)
    st = O.statistics(es, el, op.mcov_out, op.pval)
    calls, _ = O.find_events(lcp, text, suff, bwt, es, el, op, st.max_clust_length, rs.reads, O.uniform_read_offsets(*rs.reads.shape))
    g1 = rs.genome1.tobytes().decode()
    g2 = rs.genome2.tobytes().decode()
    rng = np.random.default_rng(5)

    def fasta(contigs, width=60, lower_every=7):
        out = []
        for j, (name, seq) in enumerate(contigs):
            out.append(">" + name)
            for i in range(0, len(seq), width):
                line = seq[i:i + width]
                out.append(line.lower() if (i // width + j) % lower_every == 0 else line)
        return ("\n".join(out) + "\n").encode()

    def vcf(rows, extra=()):
        lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
        lines += ["%s\t%d\t.\t%s\t%s\t.\tPASS\t." % r for r in rows]
        lines += list(extra)
        return ("\n".join(lines) + "\n").encode()

    # sample 2 has no indels in genome coordinates before a SNP only if n_indels == 0; "tiny" plants 8 short indels, whose
    # coordinates shift g2: the ALT base is read from the reads' point of view = g2 at the shifted position
    shift = np.zeros(len(g1) + 1, dtype=np.int64)
    for p, ln in rs.indels:
        shift[p:] += ln
    rows = []
    for p in rs.snp_pos.tolist():
        alt = g2[p + int(shift[p])] if 0 <= p + int(shift[p]) < len(g2) else "N"
        rows.append(("chr1", p + 1, g1[p], alt))
    indel_rows = [("chr1", p + 1, g1[p:p + 2], g1[p]) for p, ln in rs.indels if ln < 0][:3]
    rows_all = sorted(rows + indel_rows, key=lambda r: r[1])
    cases = []
    cases.append(("planted", dict(fasta=fasta([("chr1", g1)]), vcf=vcf(rows_all), calls=calls), []))
    cases.append(("k1000_l50", dict(fasta=fasta([("chr1", g1)]), vcf=vcf(rows_all), calls=calls), ["-k", "1000", "-l", "50"]))
    # two contigs with descriptions, an unknown chromosome, a SNP too close to the start, one beyond the end, position 0
    half = len(g1) // 2
    rows2 = [("ctgA first half", p, r, a) if p <= half else ("ctgB", p - half, r, a) for (_, p, r, a) in rows]
    rows2 = [(c.split()[0] if c.startswith("ctgB") else c, p, r, a) for (c, p, r, a) in rows2]
    extra = ["chrUn\t77\t.\tA\tC\t.\tPASS\t.", "ctgB\t5\t.\tA\tG\t.\tPASS\t.", "ctgB\t%d\t.\tC\tT\t.\tPASS\t." % (len(g1) - half + 50),
             "ctgB\t0\t.\tG\tT\t.\tPASS\t.", ""]
    # (istream >> reads the contig name up to the first blank: "ctgA first half" in the VCF names contig "ctgA", which the
    # FASTA calls "ctgA first half" -> not found, a WARNING per line: the reference's behaviour, kept)
    cases.append(("two_contigs", dict(fasta=fasta([("ctgA first half", g1[:half]), ("ctgB", g1[half:])]),
                                     vcf=vcf([r for r in rows2], extra), calls=calls), []))
    # damaged calls: invented SNP records (false positives), a record with two mismatching columns, no trailing newline
    recs = calls.decode().split("\n")
    fake = []
    for j in range(5):
        a = "".join(rng.choice(list("ACGT"), size=61))
        b = list(a)
        b[30] = "ACGT"[("ACGT".index(a[30]) + 1) % 4]
        if j == 4:
            b[10] = "ACGT"[("ACGT".index(a[10]) + 2) % 4]
        fake += [">SNP_higher_path_%d|P_1:30_%s/%s|5|nb_pol_1" % (900 + j, a[30], b[30]), a,
                 ">SNP_lower_path_%d|P_1:30_%s/%s|6|nb_pol_1" % (900 + j, a[30], b[30]), "".join(b)]
    damaged = "\n".join(recs[:40] + fake + recs[40:80])
    cases.append(("damaged_calls", dict(fasta=fasta([("chr1", g1)]), vcf=vcf(rows_all), calls=damaged.rstrip("\n").encode()), []))
    bad = "\n".join([">SNP_higher_path_1|P_1:30_A/C|5|nb_pol_1", "ACGTACGT", ">SNP_lower_path_1|P_1:30_A/C|5|nb_pol_1", "ACGTACG", ""])
    cases.append(("unequal_lengths", dict(fasta=fasta([("chr1", g1)]), vcf=vcf(rows_all), calls=bad.encode()), []))
    cases.append(("empty_calls", dict(fasta=fasta([("chr1", g1)]), vcf=vcf(rows_all), calls=b""), []))
    return cases


def snp_vs_vcf_golden(d):
    """stdout + exit code of the reference's snp_vs_vcf on the planted-truth inputs"""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref", "snp_vs_vcf")
    out = {}
    cases = snp_vs_vcf_inputs()
    for j, (name, files, argv) in enumerate(cases):
        paths = {}
        for k, data in files.items():
            paths[k] = os.path.join(d, f"svv{j}.{k}")
            open(paths[k], "wb").write(data)
            out[f"{j}_{k}"] = np.frombuffer(data, dtype=np.uint8)
        r = subprocess.run([ref, "-v", paths["vcf"], "-c", paths["calls"], "-f", paths["fasta"], *argv], capture_output=True, timeout=120)
        out[f"{j}_stdout"] = np.frombuffer(r.stdout, dtype=np.uint8)
        out[f"{j}_rc"] = np.int64(r.returncode)
        out[f"{j}_argv"] = np.array(argv, dtype="U16")
        out[f"{j}_name"] = np.array(name)
        print("snp_vs_vcf", name, "rc", r.returncode, r.stdout.decode().strip().splitlines()[-7:])
    out["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "snp_vs_vcf.npz"), **out)


def help_texts():
    """what the reference binaries print for -h (the drop-in CLIs print the same bytes)"""
    import subprocess
    for tool in ("ebwt2clust", "clust2snp"):
        r = subprocess.run([os.path.join(os.path.dirname(HERE), "..", "oracle", "_ref", tool), "-h"], capture_output=True)
        open(os.path.join(HERE, "help_" + tool + ".txt"), "wb").write(r.stdout)


if __name__ == "__main__":
    assert O.ref_available(), "oracle/_ref missing: run `make -C oracle ref` where /root/reference exists"
    if len(sys.argv) > 1 and sys.argv[1] == "help":  # only these fixtures
        help_texts()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "snp_vs_vcf":  # only this fixture
        d = tempfile.mkdtemp(prefix="golden_")
        try:
            snp_vs_vcf_golden(d)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        sys.exit(0)
    d = tempfile.mkdtemp(prefix="golden_")
    try:
        snp_vs_vcf_golden(d)
        phase1_fuzz(d)
        phantom_tail(d)
        phase1_fuzz_widths(d)
        for name, cfg in MICRO.items():
            micro(name, cfg, d)
        snp_text_tools(d)
        help_texts()
    finally:
        shutil.rmtree(d, ignore_errors=True)
    print("golden vectors written to", HERE)
