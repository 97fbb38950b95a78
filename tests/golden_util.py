"""Loads the committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from
the unmodified reference binaries) and rebuilds their inputs."""
import ast
import hashlib
import os

import numpy as np

from ebwt2snp_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CACHE = {}


def phase1_cases():
    z = np.load(os.path.join(GOLDEN, "phase1_fuzz.npz"))
    lo = oo = 0
    for n, k, m, ncl, nout in z["meta"]:
        yield dict(lcp=z["lcp"][lo:lo + n], bwt=z["bwt"][lo:lo + n], k=int(k), m=int(m), n_clust_out=int(ncl),
                   clusters=z["out"][oo:oo + nout].tobytes())
        lo += n
        oo += nout


def phase1_width_cases():
    z = np.load(os.path.join(GOLDEN, "phase1_fuzz_widths.npz"))
    lo = oo = 0
    for n, k, m, ncl, nout, x, y, zz in z["meta"]:
        yield dict(lcp=z["lcp"][lo:lo + n], bwt=z["bwt"][lo:lo + n], k=int(k), m=int(m), n_clust_out=int(ncl),
                   clusters=z["out"][oo:oo + nout].tobytes(), x=int(x), y=int(y), z=int(zz))
        lo += n
        oo += nout


def phantom_tail_cases():
    """hand-made indexes whose last cluster contains the post-EOF phantom record; `readouts` = [(n1, n_candidates)]
    as printed by the reference with -L 1 -R 1 (tests/golden/make_golden.py:phantom_tail)"""
    z = np.load(os.path.join(GOLDEN, "phantom_tail.npz"))
    rows = z["rows"]
    for ci in range(int(z["n_cases"])):
        mine = rows[rows[:, 0] == ci]
        x, y, zz, bcr = (int(v) for v in mine[0, 1:5])
        yield dict(ci=ci, x=x, y=y, z=zz, bcr=bool(bcr), lcp=z[f"c{ci}_lcp"].astype(np.uint32), text=z[f"c{ci}_text"].astype(np.uint32),
                   suff=z[f"c{ci}_suff"].astype(np.uint32), bwt=z[f"c{ci}_bwt"], clusters=z[f"c{ci}_clusters"].tobytes(),
                   readouts=[(int(r[5]), int(r[6])) for r in mine])


def unpack2(packed, shape):
    R, L = int(shape[0]), int(shape[1])
    c = np.stack([(packed >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)[: R * L]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[c].reshape(R, L)


def micro(name):
    """dict: reads, nreads1, egsa (numpy u32/u8 arrays), k, m, clusters bytes, n_clust_out, variants[]"""
    if name in _CACHE:
        return _CACHE[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    reads = unpack2(z["reads2bit"], z["shape"])
    e = synth.build_egsa(reads)
    eg = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in e.items()}
    for f in ("lcp", "text", "suff"):
        eg[f] = eg[f].view(np.uint32)
    rec = synth.gesa_records(eg)
    assert hashlib.sha256(rec.tobytes()).hexdigest() == str(z["gesa_sha256"]), "regenerated EGSA differs from the golden input"
    variants = []
    for vi in range(int(z["n_variants"])):
        variants.append(dict(args=str(z[f"v{vi}_args"]).split(), rc=int(z[f"v{vi}_rc"]),
                             allowed=tuple(int(x) for x in z[f"v{vi}_allowed"]), ncand=int(z[f"v{vi}_ncand"]),
                             snp=z[f"v{vi}_snp"].tobytes()))
    layouts = []
    for j in range(int(z["n_layouts"])):
        x, y, zz, bcr = (int(v) for v in z[f"lay{j}_spec"])
        layouts.append(dict(x=x, y=y, z=zz, bcr=bool(bcr), n_clust_out=int(z[f"lay{j}_nclust"]),
                            clusters=z[f"lay{j}_clusters"].tobytes(), rc=int(z[f"lay{j}_rc"]),
                            allowed=tuple(int(v) for v in z[f"lay{j}_allowed"]), ncand=int(z[f"lay{j}_ncand"]),
                            snp=z[f"lay{j}_snp"].tobytes()))
    d = dict(name=name, reads=reads, nreads1=int(z["nreads1"]), egsa=eg, gesa=rec, k=int(z["k"]), m=int(z["m"]), layouts=layouts,
             clusters=z["clusters"].tobytes(), n_clust_out=int(z["n_clust_out"]), variants=variants,
             gen=ast.literal_eval(str(z["gen"])))
    _CACHE[name] = d
    return d


ARG_TO_PARAM = {"-m": "mcov_out", "-c": "consensus_reads", "-g": "max_gap", "-L": "k_left", "-R": "k_right",
                "-e": "max_err", "-p": "pval", "-v": None}


def params_kw(args):
    kw = {}
    for flag, val in zip(args[0::2], args[1::2]):
        name = ARG_TO_PARAM[flag]
        if name is None:
            continue  # -v is parsed and ignored by the reference
        kw[name] = float(val) if name == "pval" else int(val)
    return kw
