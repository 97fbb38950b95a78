"""Seeded synthetic two-sample read sets and their EGSA (eBWT + LCP + generalized SA).

This is DATA PREPARATION, not the hot path: it stands in for the external `egsa` / BCR tools
the reference expects to have been run beforehand (ref:README.md:46-60, ref:pipeline.sh:98-109).
It is used by the tests, by `bench.py` (outside the timed region) and by the golden-fixture
script.  Conventions (last row; the reference pins none of them):

  * one record per suffix of every read INCLUDING the terminator suffix -> n = R * (L + 1);
  * order: `$` < A < C < G < T; equal strings (incl. terminator) ordered by read id;
  * `lcp[i]` = common prefix with record i-1, never extending over a terminator; lcp[0] = 0;
  * `text` = read id, `suff` = offset of the suffix in the read (L for the terminator suffix);
  * `bwt`  = preceding character, `$` (0x24) for whole-read suffixes;
  * alphabet strictly ACGT (an `N` makes the reference non-deterministic: ref:include.hpp:273).

The random choices are made with numpy on the host (so the same seed gives the same reads
whether the suffix sort then runs on the CPU or on a GPU); the heavy part (key building,
stable radix sorts, LCP) is plain torch on whichever device is asked for.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[ord("A")], _COMP[ord("C")], _COMP[ord("G")], _COMP[ord("T")] = ord("T"), ord("G"), ord("C"), ord("A")
TERMINATOR = 0x24  # '$'
SYMS_PER_WORD = 21  # 3-bit symbols in a 63-bit sort key


@dataclass
class ReadSet:
    reads: np.ndarray  # (R, L) uint8 ASCII, file order: s1 [, s1-RC], s2 [, s2-RC] (ref:pipeline.sh:92-94)
    nreads1: int  # value for clust2snp -n: records belonging to sample 1
    genome1: np.ndarray
    genome2: np.ndarray
    snp_pos: np.ndarray  # planted SNP positions in genome1 coordinates
    indels: list  # (pos, +len inserted in sample 2 | -len deleted in sample 2)


# named configurations of BASELINE.json (sizes: SURVEY.md §8(d))
CONFIGS = {
    # name: genome length, reads per sample, read length, #SNPs, #indels, reverse complements
    "tiny": dict(G=20_000, reads_per_sample=4_000, L=100, n_snps=40, n_indels=8, rc=True),
    "small": dict(G=100_000, reads_per_sample=20_000, L=100, n_snps=100, n_indels=20, rc=True),
    "C1": dict(G=1_000_000, reads_per_sample=50_000, L=100, n_snps=1_000, n_indels=0, rc=False),
    "C2": dict(G=4_600_000, reads_per_sample=1_380_000, L=100, n_snps=4_600, n_indels=0, rc=True),
    "C3": dict(G=64_000_000, reads_per_sample=19_200_000, L=100, n_snps=64_000, n_indels=6_400, rc=False),
    # C4 (3 Gbp, 2 x 20x: 1.21e11 positions, 1.58 TB of .gesa records) is never generated in one piece: bench.py streams it as
    # tiles of the C2 index with shifted read ids (every tile starts with lcp = 0, no cluster spans tiles)
    "C4": dict(G=3_000_000_000, reads_per_sample=600_000_000, L=100, n_snps=3_000_000, n_indels=0, rc=False),
}


def make_read_set(G, reads_per_sample, L=100, n_snps=0, n_indels=0, rc=False, seed=1, sample_chunk=1 << 20) -> ReadSet:
    rng = np.random.default_rng(seed)
    g1 = BASES[rng.integers(0, 4, size=G)]
    g2 = g1.copy()
    snp_pos = np.sort(rng.choice(G, size=n_snps, replace=False)) if n_snps else np.zeros(0, np.int64)
    if n_snps:
        # a different base: rotate by 1..3 in ACGT order
        code = np.searchsorted(BASES, g2[snp_pos])
        g2[snp_pos] = BASES[(code + rng.integers(1, 4, size=n_snps)) % 4]
    indels = []
    if n_indels:
        pos = np.sort(rng.choice(np.arange(L, G - L), size=n_indels, replace=False))
        pieces, last = [], 0
        for p in pos:
            ln = int(rng.integers(1, 4))
            if rng.integers(0, 2):  # insertion in sample 2
                pieces += [g2[last:p], BASES[rng.integers(0, 4, size=ln)]]
                last = p
                indels.append((int(p), ln))
            else:  # deletion in sample 2
                pieces.append(g2[last:p])
                last = p + ln
                indels.append((int(p), -ln))
        pieces.append(g2[last:])
        g2 = np.concatenate(pieces)

    def sample(g):
        st = rng.integers(0, len(g) - L + 1, size=reads_per_sample)
        out = np.empty((reads_per_sample, L), dtype=np.uint8)
        cols = np.arange(L)[None, :]
        for lo in range(0, reads_per_sample, sample_chunk):  # bounded index temporaries (C3: 19.2 M reads per sample)
            hi = min(reads_per_sample, lo + sample_chunk)
            out[lo:hi] = g[st[lo:hi, None] + cols]
        return out

    r1, r2 = sample(g1), sample(g2)
    if rc:
        parts = [r1, _COMP[r1[:, ::-1]], r2, _COMP[r2[:, ::-1]]]
        nreads1 = 2 * reads_per_sample
    else:
        parts = [r1, r2]
        nreads1 = reads_per_sample
    return ReadSet(np.ascontiguousarray(np.concatenate(parts)), nreads1, g1, g2, snp_pos, indels)


def make_config(name: str, seed: int = 1, scale: float = 1.0) -> ReadSet:
    """Read set of a named BASELINE config; `scale` shrinks genome and read count together
    (same coverage, same read length) for bounded CPU-baseline samples."""
    c = dict(CONFIGS[name])
    if scale != 1.0:
        for key in ("G", "reads_per_sample", "n_snps", "n_indels"):
            c[key] = max(int(c[key] * scale), 0 if key.startswith("n_") else 1000)
    return make_read_set(seed=seed, **c)


# ------------------------------------------------------------------------------------------
# EGSA construction (suffix sort of the read collection)
# ------------------------------------------------------------------------------------------

def _codes(reads_t: torch.Tensor) -> torch.Tensor:
    """ASCII (R, L) -> flat 3-bit codes with a terminator column: $=0 A=1 C=2 G=3 T=4."""
    lut = torch.zeros(256, dtype=torch.uint8, device=reads_t.device)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i + 1
        lut[ch + 32] = i + 1
    R, L = reads_t.shape
    codes = torch.zeros((R, L + 1), dtype=torch.uint8, device=reads_t.device)
    codes[:, :L] = lut[reads_t.long()]
    flat = torch.cat([codes.reshape(-1), torch.zeros(SYMS_PER_WORD * 8, dtype=torch.uint8, device=reads_t.device)])
    return flat


def _key_words(flat, n, L, W):
    """63-bit keys in natural suffix order idx = r*(L+1)+p: word w holds symbols [21w, 21w+21)
    of the suffix, `$`(=0)-padded past its terminator.  Contiguous slices only (no gathers)."""
    p = torch.arange(n, dtype=torch.int64, device=flat.device) % (L + 1)
    keys = []
    for w in range(W):
        key = torch.zeros(n, dtype=torch.int64, device=flat.device)
        for j in range(SYMS_PER_WORD):
            off = w * SYMS_PER_WORD + j
            if off > L:
                break
            sym = flat[off : off + n].long()
            sym.masked_fill_(p + off > L, 0)
            key |= sym << (3 * (SYMS_PER_WORD - 1 - j))
        keys.append(key)
    return keys


def _bit_length(x: torch.Tensor) -> torch.Tensor:
    """bit length of non-negative int64 (0 -> 0), exact (halves go through float64 frexp)."""
    hi, lo = x >> 32, x & 0xFFFFFFFF
    bl_hi = torch.frexp(hi.double())[1].long()
    bl_lo = torch.frexp(lo.double())[1].long()
    return torch.where(hi > 0, bl_hi + 32, torch.where(lo > 0, bl_lo, torch.zeros_like(lo)))


def build_egsa(reads: np.ndarray, device="cpu", chunk: int = 1 << 27):
    """Suffix-sort the collection; returns dict of torch tensors on `device`:
    lcp (int32 holding u32), text (int32), suff (int32), bwt (uint8), all of length R*(L+1)."""
    dev = torch.device(device)
    reads_t = torch.from_numpy(reads).to(dev)
    R, L = reads_t.shape
    n = R * (L + 1)
    W = (L + 1 + SYMS_PER_WORD - 1) // SYMS_PER_WORD
    flat = _codes(reads_t)
    keys = _key_words(flat, n, L, W)
    perm = torch.arange(n, dtype=torch.int64, device=dev)
    # LSD stable sorts, least significant word first; initial order (r, p) breaks ties by read id
    for w in reversed(range(W)):
        order = torch.sort(keys[w][perm], stable=True)[1]
        perm = perm[order]
        del order
    text = (perm // (L + 1)).to(torch.int32)
    suff = (perm % (L + 1)).to(torch.int32)
    # BWT: preceding char, '$' for whole-read suffixes
    flat_reads = reads_t.reshape(-1)
    prev = (perm // (L + 1)) * L + (perm % (L + 1)) - 1
    bwt = torch.where(suff > 0, flat_reads[prev.clamp_(min=0)], torch.full_like(flat_reads[:1], TERMINATOR).expand(n))
    del prev
    # LCP with the previous record
    lcp = torch.zeros(n, dtype=torch.int64, device=dev)
    for lo in range(1, n, chunk):
        hi = min(n, lo + chunk)
        a, b = perm[lo - 1 : hi - 1], perm[lo:hi]
        pa, pb = a % (L + 1), b % (L + 1)
        res = torch.minimum(L - pa, L - pb)
        undecided = torch.ones_like(res, dtype=torch.bool)
        for w in range(W):
            x = keys[w][a] ^ keys[w][b]
            first = undecided & (x != 0)
            # symbol j occupies bits [3*(20-j), 3*(20-j)+3): j = 20 - (bit_length-1)//3
            j = (SYMS_PER_WORD - 1) - (_bit_length(x) - 1) // 3
            res = torch.where(first, torch.minimum(res, w * SYMS_PER_WORD + j), res)
            undecided &= ~first
            if not bool(undecided.any()):
                break
        lcp[lo:hi] = res
    return dict(lcp=lcp.to(torch.int32), text=text, suff=suff, bwt=bwt.contiguous(), n=n, L=L, R=R)


# ------------------------------------------------------------------------------------------
# file formats (ref:include.hpp:124-155 for .gesa; FASTA as read by ref:clust2snp.cpp:147-212)
# ------------------------------------------------------------------------------------------

GESA_DTYPE_444 = np.dtype([("text", "<u4"), ("suff", "<u4"), ("lcp", "<u4"), ("bwt", "u1")])


def _np(a):
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def gesa_records(egsa, x=4, y=4, z=4) -> np.ndarray:
    """Interleave into the on-disk record order text(y) suff(z) lcp(x) bwt(1), little endian."""
    dt = np.dtype([("text", f"<u{y}"), ("suff", f"<u{z}"), ("lcp", f"<u{x}"), ("bwt", "u1")])
    n = int(egsa["n"])
    rec = np.empty(n, dtype=dt)
    for name in ("text", "suff", "lcp"):
        rec[name] = _np(egsa[name]).view(np.uint32)  # narrowing cast truncates like the file would
    rec["bwt"] = _np(egsa["bwt"])
    return rec


def write_gesa(path, egsa, x=4, y=4, z=4):
    gesa_records(egsa, x, y, z).tofile(path)


def write_bcr(prefix, egsa, x=4, y=4, z=4):
    """BCR triple: .out (bwt), .out.lcp (lcp), .out.pairSA (suff(z) then text(y)) ref:include.hpp:157-188."""
    _np(egsa["bwt"]).tofile(prefix + ".out")
    _np(egsa["lcp"]).view(np.uint32).astype(f"<u{x}").tofile(prefix + ".out.lcp")
    dt = np.dtype([("suff", f"<u{z}"), ("text", f"<u{y}")])
    rec = np.empty(int(egsa["n"]), dtype=dt)
    rec["suff"] = _np(egsa["suff"]).view(np.uint32)
    rec["text"] = _np(egsa["text"]).view(np.uint32)
    rec.tofile(prefix + ".out.pairSA")


def write_fasta(path, reads: np.ndarray):
    R, L = reads.shape
    hdr = np.char.add(np.char.add(">r", np.arange(R).astype(str)), "\n").astype("S")
    with open(path, "wb") as f:
        # vectorised: header bytes are ragged, so write in blocks
        blk = 1 << 16
        nl = np.full((1,), 10, dtype=np.uint8)
        for lo in range(0, R, blk):
            hi = min(R, lo + blk)
            out = bytearray()
            for i in range(lo, hi):
                out += hdr[i]
                out += reads[i].tobytes()
                out += b"\n"
            f.write(out)
        del nl


def write_fasta_fixed(path, reads: np.ndarray):
    """Same reads, fixed-width headers (`>` + 9 digits): the whole file is one 2-D byte array, written without a Python
    loop per read (C2 / C3-size read sets).  The tools only look at the first byte of a header line."""
    R, L = reads.shape
    out = np.empty((R, 11 + L + 1), dtype=np.uint8)
    out[:, 0] = ord(">")
    ids = np.arange(R, dtype=np.int64)
    for j in range(9):
        out[:, 9 - j] = (ids // 10 ** j % 10 + ord("0")).astype(np.uint8)
    out[:, 10] = 10
    out[:, 11:11 + L] = reads
    out[:, 11 + L] = 10
    out.tofile(path)


def write_dataset(dirpath, rs: ReadSet, egsa, name="ALL.fasta", x=4, y=4, z=4, bcr=False, fixed_headers=False):
    os.makedirs(dirpath, exist_ok=True)
    fasta = os.path.join(dirpath, name)
    (write_fasta_fixed if fixed_headers else write_fasta)(fasta, rs.reads)
    if bcr:
        write_bcr(fasta, egsa, x, y, z)
    else:
        write_gesa(fasta + ".gesa", egsa, x, y, z)
    return fasta
